"""GEMM timing free of host launch gaps: R launches captured in one CUDA graph, bracketed by events (warm L2 across
launches unless --flush interleaves a 256 MB memset whose time is measured separately and subtracted)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=10); ap.add_argument("--flush", action="store_true")
ap.add_argument("--only", default=""); ap.add_argument("--bn", type=int, default=0); ap.add_argument("--cublas", action="store_true", help="also time torch.matmul (cuBLAS bf16, no epilogue) on the same shape as a library bar")
a = ap.parse_args()
F32, BF = torch.float32, torch.bfloat16
SHAPES = [
    ("s0.qkv", 200704, 288, 96, 0, BF, None), ("s0.proj", 200704, 96, 96, 0, F32, F32), ("s0.fc1", 200704, 384, 96, 1, BF, None), ("s0.fc2", 200704, 96, 384, 0, F32, F32),
    ("s1.qkv", 50176, 576, 192, 0, BF, None), ("s1.proj", 50176, 192, 192, 0, F32, F32), ("s1.fc1", 50176, 768, 192, 1, BF, None), ("s1.fc2", 50176, 192, 768, 0, F32, F32),
    ("s2.qkv", 12544, 1152, 384, 0, BF, None), ("s2.proj", 12544, 384, 384, 0, F32, F32), ("s2.fc1", 12544, 1536, 384, 1, BF, None), ("s2.fc2", 12544, 384, 1536, 0, F32, F32),
    ("s3.qkv", 3136, 2304, 768, 0, BF, None), ("s3.proj", 3136, 768, 768, 0, F32, F32), ("s3.fc1", 3136, 3072, 768, 1, BF, None), ("s3.fc2", 3136, 768, 3072, 0, F32, F32),
    ("bert.qkv", 8384, 2304, 768, 0, BF, None), ("bert.ao", 8384, 768, 768, 0, F32, F32), ("bert.fi", 8384, 3072, 768, 1, BF, None), ("bert.fo", 8384, 768, 3072, 0, F32, F32),
]
COUNT = {"s0": 2, "s1": 2, "s2": 18, "s3": 2, "bert": 12}
_lib.ensure_init()
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
t_flush = graph_time(lambda: flush.zero_(), a.reps) if a.flush else 0.0
total = 0.0
for name, M, N, K, act, od, rd in SHAPES:
    if a.only and a.only not in name: continue
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, device="cuda"); r = None if rd is None else torch.randn(M, N, device="cuda").to(rd)
    out = r if (rd == od and r is not None) else torch.empty(M, N, device="cuda", dtype=od)   # in-place residual as in the model
    def fn():
        if a.flush: flush.zero_()
        ops.linear(x, w, b, act=act, residual=r, out=out, block_n=a.bn)
    ms = graph_time(fn, a.reps) - t_flush
    total += COUNT[name.split(".")[0]] * ms
    byts = M * K * 2 + N * K * 2 + M * N * (od.itemsize + (rd.itemsize if rd else 0))
    extra = ""
    if a.cublas:
        wt = w.t().contiguous().t(); ob = torch.empty(M, N, device="cuda", dtype=BF)
        def fn2():
            if a.flush: flush.zero_()
            torch.matmul(x, wt.t() if False else w.t(), out=ob)
        ms2 = graph_time(fn2, a.reps) - t_flush
        extra = f"  | cuBLAS bf16 (no epilogue) {ms2*1e3:7.1f} us {2*M*N*K/ms2/1e9:7.1f} TF/s"
    print(f"{name:9s} {M:7d}x{N:5d}x{K:5d} {ms*1e3:8.1f} us {2*M*N*K/ms/1e9:7.1f} TF/s  {byts/ms/1e6:7.0f} GB/s(algorithmic){extra}", flush=True)
print(f"sum over the step: {total:.3f} ms (flush {t_flush*1e3:.1f} us subtracted)" )
