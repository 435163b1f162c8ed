"""Pipeline trace of CTA 0 of gemm_tc_kernel (clock64 stamps through the mvlt_debug_gemm_trace hook)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops, _lib
lib = _lib.ensure_init()
lib.mvlt_debug_gemm_trace.argtypes = [ctypes.c_void_p]
F32, BF = torch.float32, torch.bfloat16
CASES = {"s2.qkv": (12544, 1152, 384, 0, BF, None), "s2.proj": (12544, 384, 384, 0, F32, F32), "s2.fc1": (12544, 1536, 384, 1, BF, None),
         "s2.fc2": (12544, 384, 1536, 0, F32, F32), "bert.qkv": (8384, 2304, 768, 0, BF, None), "bert.ao": (8384, 768, 768, 0, F32, F32),
         "bert.fi": (8384, 3072, 768, 1, BF, None), "bert.fo": (8384, 768, 3072, 0, F32, F32)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in (sys.argv[1:] or list(CASES)):
    M, N, K, act, od, rd = CASES[name]
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, device="cuda"); r = None if rd is None else torch.randn(M, N, device="cuda").to(rd)
    out = r if r is not None else torch.empty(M, N, device="cuda", dtype=od)
    for _ in range(3): ops.linear(x, w, b, act=act, residual=r, out=out)
    for cold in (False, True):
        tr = torch.zeros(512, dtype=torch.int64, device="cuda")
        if cold: flush.zero_()
        lib.mvlt_debug_gemm_trace(tr.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(x, w, b, act=act, residual=r, out=out); e1.record()
        torch.cuda.synchronize()
        lib.mvlt_debug_gemm_trace(None)
        t = tr.cpu().tolist(); t0 = t[0]
        rel = lambda i: (t[i] - t0) if t[i] else None
        print(f"== {name} {M}x{N}x{K} {'cold' if cold else 'warm'}: event {e0.elapsed_time(e1)*1e3:.1f} us | setup done {rel(1)} pdl {rel(2)} epi end {rel(3)} exit {rel(4)} clk")
        if not cold:
            print("  epi warp 0 chunk steps: start | store-drain+prefetch | acquired | tmem ld done | math done | st.shared done | fence+syncwarp | store issued   (deltas)")
            for k in range(12):
                v = [t[256 + 8 * k + i] for i in range(8)]
                if not v[0]: break
                print(f"   k={k:2d} @{v[0]-t0:6d}: " + " ".join(f"{(v[i]-v[i-1]) if v[i] and v[i-1] else -1:5d}" for i in range(1, 8)))
        for it in range(16):
            if not t[16 + 4 * it]: break
            print(f"  tile {it}: mma acquires acc {rel(16+4*it)}, first operands {rel(17+4*it)}, last mma issued {rel(18+4*it)} | epi0 acquires {rel(128+4*it)}, releases {rel(129+4*it)}")
