"""Quick on-box check + timing of the fused Swin block-tail kernel against the unfused chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from medical_vision_langauge_transformer_b200 import ops

def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()
def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
def graph_time(fn, n=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); g.replay(); e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

for (M, C) in [(256, 96), (300, 96), (200704, 96), (12544, 384), (50176, 192)]:
    for with_proj in (True, False):
        x = rnd(M, C, seed=1); o = rnd(M, C, seed=2).bfloat16()
        wp, bp = rnd(C, C, seed=3, scale=C ** -0.5).bfloat16(), rnd(C, seed=4, scale=0.1)
        g, b = 1 + rnd(C, seed=5, scale=0.1), rnd(C, seed=6, scale=0.1)
        w1, b1 = rnd(4 * C, C, seed=7, scale=C ** -0.5).bfloat16(), rnd(4 * C, seed=8, scale=0.1)
        w2, b2 = rnd(C, 4 * C, seed=9, scale=(4 * C) ** -0.5).bfloat16(), rnd(C, seed=10, scale=0.1)
        x1 = x + o.float() @ wp.float().t() + bp if with_proj else x.clone()
        a = F.layer_norm(x1, (C,), g, b, 1e-5).bfloat16().float()
        h = F.gelu(a @ w1.float().t() + b1).bfloat16().float()
        ref = x1 + h @ w2.float().t() + b2
        xf = x.clone()
        ops.swin_block_tail(xf, o if with_proj else None, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
        torch.cuda.synchronize()
        err = relerr(xf, ref)
        bad = (~torch.isfinite(xf)).sum().item()
        line = f"tail M={M} C={C} proj={with_proj}: relerr {err:.2e} nonfinite {bad}"
        if err > 1e-2:
            d = (xf - ref).abs()
            rows = (d.amax(1) > 0.05 * ref.abs().max()).nonzero().flatten()
            cols = (d.amax(0) > 0.05 * ref.abs().max()).nonzero().flatten()
            line += f"  bad rows {rows[:8].tolist()}..({rows.numel()}) bad cols {cols[:8].tolist()}..({cols.numel()})"
        if M >= 12544:
            xs = [x.clone() for _ in range(4)]; k = [0]
            def fused():
                k[0] = (k[0] + 1) % 4
                ops.swin_block_tail(xs[k[0]], o if with_proj else None, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
            def chain():
                k[0] = (k[0] + 1) % 4
                xc = xs[k[0]]
                if with_proj: ops.linear(o, wp, bp, residual=xc, out=xc)
                an = ops.layernorm(xc, g, b, 1e-5, torch.bfloat16)
                hn = ops.linear(an, w1, b1, act=ops.ACT_GELU)
                ops.linear(hn, w2, b2, residual=xc, out=xc)
            line += f"   fused {graph_time(fused):.1f} us   chain {graph_time(chain):.1f} us (in-graph, back to back)"
        print(line, flush=True)
