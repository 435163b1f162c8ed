"""Same GEMM shape under the epilogue variants (no residual fp32 / bf16 out, in-place fp32 reduce-add, out-of-place fp32
residual, bf16 out + bf16 residual): separates the cost of the residual handling from the MMA time.  L2 flushed between."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
_lib.ensure_init()
F32, BF = torch.float32, torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def graph_time(fn, reps):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps): fn()
        g.replay(); st.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); st.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] / reps
t_flush = graph_time(lambda: flush.zero_(), a.reps)
for name, M, N, K in [("bert.ao", 8384, 768, 768), ("bert.fo", 8384, 768, 3072), ("s2.proj", 12544, 384, 384), ("s2.fc2", 12544, 384, 1536),
                      ("s0.proj", 200704, 96, 96), ("s1.proj", 50176, 192, 192)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16(); b = torch.randn(N, device="cuda")
    variants = {
        "f32 out": dict(out=torch.empty(M, N, device="cuda", dtype=F32)),
        "bf16 out": dict(out=torch.empty(M, N, device="cuda", dtype=BF)),
        "f32 in-place +=": None,
        "f32 out-of-place res": dict(out=torch.empty(M, N, device="cuda", dtype=F32), residual=torch.randn(M, N, device="cuda")),
        "bf16 out + bf16 res": dict(out=torch.empty(M, N, device="cuda", dtype=BF), residual=torch.randn(M, N, device="cuda").bfloat16()),
    }
    r = torch.randn(M, N, device="cuda")
    variants["f32 in-place +="] = dict(out=r, residual=r)
    line = f"{name:8s} {M}x{N}x{K}:"
    for vn, kw in variants.items():
        def fn():
            flush.zero_(); ops.linear(x, w, b, **kw)
        ms = graph_time(fn, a.reps) - t_flush
        line += f"  {vn} {ms*1e3:6.1f} us"
    print(line, flush=True)
