"""LayerNorm row kernel timing, warm L2 (the producer GEMM has just written the rows) and cold (256 MB flush in between)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def graph_time(fn, reps=10):
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
    ts.sort(); return ts[len(ts) // 2] / reps * 1e3
tf = graph_time(lambda: flush.zero_())
for name, rows, C, od, copy in [("bert post-LN", 8384, 768, torch.float32, True), ("swin s2 LN", 12544, 384, torch.bfloat16, False),
                                ("swin s1 LN", 50176, 192, torch.bfloat16, False), ("swin s0 LN", 200704, 96, torch.bfloat16, False)]:
    x = torch.randn(rows, C, device="cuda"); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    out = x if od == torch.float32 else torch.empty(rows, C, device="cuda", dtype=od)
    fn = lambda: ops.layernorm(x, g, b, 1e-5, od, out=out, bf16_copy=copy)
    warm = graph_time(fn)
    cold = graph_time(lambda: (flush.zero_(), fn())) - tf
    byts = rows * C * (4 + out.element_size() + (2 if copy else 0))
    print(f"{name:14s} rows {rows:6d} C {C:4d}: warm {warm:6.1f} us ({byts / warm / 1e6:6.0f} GB/s)   cold {cold:6.1f} us ({byts / cold / 1e6:6.0f} GB/s)")
