"""Quick on-box check of the tcgen05 attention kernels (run before the full test-suite): prints errors per case instead
of stopping at the first assertion."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import ops

def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()

def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "window"):
    for (B, H, C, heads, shift) in [(2, 14, 384, 12, 0), (2, 14, 384, 12, 3), (1, 7, 768, 24, 0), (2, 56, 96, 3, 3), (64, 14, 384, 12, 3),
                                    (64, 56, 96, 3, 0), (64, 28, 192, 6, 3), (64, 7, 768, 24, 0)]:
        qkv = rnd(B * H * H, 3 * C, seed=B + H).bfloat16()
        relb = torch.zeros(heads, 64, 64, device="cuda"); relb[:, :49, :49] = rnd(heads, 49, 49, seed=7, scale=0.5)
        ref = ops.window_attention(qkv.float(), relb, B, H, H, C, heads, 7, shift, 32 ** -0.5)
        perm = ops.window_major_index(B, H, H, 7, shift, device="cuda")
        qkv_wm = torch.empty_like(qkv); qkv_wm[perm] = qkv
        table = ops.window_bias_table(relb, shift)
        frag = ops.window_bias_fragments(relb, shift, 32 ** -0.5)
        try:
            out = ops.window_attention_tc(qkv_wm, table, B, H, H, C, heads, 7, shift, 32 ** -0.5)
            torch.cuda.synchronize()
            err = relerr(out, ref)
            bad = (~torch.isfinite(out.float())).sum().item()
            t_new = timeit(lambda: ops.window_attention_tc(qkv_wm, table, B, H, H, C, heads, 7, shift, 32 ** -0.5))
            t_old = timeit(lambda: ops.window_attention(qkv, frag, B, H, H, C, heads, 7, shift, 32 ** -0.5))
            gb = B * H * H * C * 8 / 1e9
            print(f"window B={B} H={H} C={C} shift={shift}: relerr {err:.2e} nonfinite {bad}  tc {t_new:.1f} us ({gb / t_new * 1e6:.0f} GB/s)  warp {t_old:.1f} us", flush=True)
            if err > 1e-2:
                d = (out.float() - ref).abs().view(B, H, H, heads, 32).amax(-1)
                print("   worst (b,h,x,head):", [tuple(int(v) for v in idx) for idx in (d > 0.05 * ref.abs().max()).nonzero()[:12]])
        except Exception as e:
            print(f"window B={B} H={H} C={C} shift={shift}: EXCEPTION {e}", flush=True)
            raise
if which in ("all", "joint"):
    for (B, S, s2s) in [(3, 131, False), (3, 131, True), (7, 74, False), (5, 81, True), (64, 131, False), (64, 74, False)]:
        heads, D = 12, 768
        qkv = rnd(B * S, 3 * D, seed=S + B).bfloat16()
        g = torch.Generator().manual_seed(S)
        valid = torch.randint(52, S + 1, (B,), generator=g)
        kmask = torch.where(torch.arange(S)[None] < valid[:, None], 0.0, -10000.0).float().cuda().contiguous()
        ref = ops.joint_attention(qkv.float(), kmask, B, S, heads, s2s, 50)
        out = ops.joint_attention(qkv, kmask, B, S, heads, s2s, 50, impl="tc")
        torch.cuda.synchronize()
        err = relerr(out, ref)
        bad = (~torch.isfinite(out.float())).sum().item()
        t_new = timeit(lambda: ops.joint_attention(qkv, kmask, B, S, heads, s2s, 50, impl="tc"))
        t_old = timeit(lambda: ops.joint_attention(qkv, kmask, B, S, heads, s2s, 50, impl="warp"))
        gb = B * S * D * 8 / 1e9
        print(f"joint B={B} S={S} seq2seq={s2s}: relerr {err:.2e} nonfinite {bad}  tc {t_new:.1f} us ({gb / t_new * 1e6:.0f} GB/s)  warp {t_old:.1f} us", flush=True)
        if err > 1e-2:
            d = (out.float() - ref).abs().view(B, S, heads, 64).amax(-1)
            print("   worst (b,s,head):", [tuple(int(v) for v in idx) for idx in (d > 0.05 * ref.abs().max()).nonzero()[:12]])
