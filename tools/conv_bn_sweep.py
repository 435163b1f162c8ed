"""Tile-width sweep for the ResNet layer-3 call sites (M = 12544 = 49 row tiles of 256 on 74 CTA pairs): us per launch, L2 flushed."""
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
B, H = 64, 14
x = torch.randn(B * H * H, 256, device="cuda").bfloat16(); w = (torch.randn(256, 9 * 256, device="cuda") / 48).bfloat16(); b = torch.randn(256, device="cuda")
for bn in (0, 64, 96, 128, 192, 256):
    t = graph_time(lambda: (flush.zero_(), ops.conv2d_nhwc(x, w, b, B, H, H, 3, 3, 1, 1, act=ops.ACT_RELU, block_n=bn))) - tf
    print(f"conv3x3 14x14 C256->256  block_n {bn:3d}: {t:6.1f} us")
x2 = torch.randn(B * H * H, 1024, device="cuda").bfloat16(); w2 = (torch.randn(256, 1024, device="cuda") / 32).bfloat16()
for bn in (0, 64, 96, 128, 192, 256):
    t = graph_time(lambda: (flush.zero_(), ops.linear(x2, w2, b, act=ops.ACT_RELU, out_dtype=torch.bfloat16, block_n=bn))) - tf
    print(f"conv1x1 1024->256        block_n {bn:3d}: {t:6.1f} us")
x3 = torch.randn(B * H * H, 256, device="cuda").bfloat16(); w3 = (torch.randn(1024, 256, device="cuda") / 16).bfloat16(); b3 = torch.randn(1024, device="cuda")
r3 = torch.randn(B * H * H, 1024, device="cuda").bfloat16()
for bn in (0, 64, 128, 192, 256):
    t = graph_time(lambda: (flush.zero_(), ops.linear(x3, w3, b3, act=ops.ACT_RELU, residual=r3, out_dtype=torch.bfloat16, block_n=bn))) - tf
    print(f"conv1x1 256->1024 + id   block_n {bn:3d}: {t:6.1f} us")
