"""Fused Swin MLP kernel vs the unfused LN -> GEMM(GELU) -> GEMM(+res) chain at the batch-64 stage shapes: R calls in one
CUDA graph between events, optional 256 MB L2 flush before every call (its own time measured and subtracted)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=10); ap.add_argument("--flush", action="store_true")
a = ap.parse_args()
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
total_f = total_u = 0.0
for name, M, C, count in (("s0", 200704, 96, 2), ("s1", 50176, 192, 2), ("s2", 12544, 384, 18)):
    x = torch.randn(M, C, device="cuda"); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    w1 = (torch.randn(4 * C, C, device="cuda") * C ** -0.5).bfloat16(); b1 = torch.randn(4 * C, device="cuda") * 0.1
    w2 = (torch.randn(C, 4 * C, device="cuda") * (4 * C) ** -0.5).bfloat16() * 0.1; b2 = torch.randn(C, device="cuda") * 0.1
    def fused():
        if a.flush: flush.zero_()
        ops.swin_mlp(x, g, b, 1e-5, w1, b1, w2, b2)
    def chain():
        if a.flush: flush.zero_()
        t = ops.layernorm(x, g, b, 1e-5, torch.bfloat16)
        h = ops.linear(t, w1, b1, act=ops.ACT_GELU)
        ops.linear(h, w2, b2, residual=x, out=x)
    tf = graph_time(fused, a.reps) - t_flush; tu = graph_time(chain, a.reps) - t_flush
    total_f += count * tf; total_u += count * tu
    flops = 2 * M * C * 4 * C * 2; byts = M * C * 8
    print(f"{name} M={M:6d} C={C:3d}: fused {tf*1e3:7.1f} us ({flops/tf/1e9:6.0f} TF/s, {byts/tf/1e6:5.0f} GB/s algorithmic)   "
          f"unfused chain {tu*1e3:7.1f} us   x{tu/tf:.2f}", flush=True)
print(f"per step: fused {total_f:.3f} ms vs unfused {total_u:.3f} ms (flush {t_flush*1e3:.1f} us subtracted)")
