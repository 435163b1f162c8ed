"""Launch one ResNet convolution call site a few times (for ncu --set full): implicit-GEMM conv via im2col-mode TMA."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops
ap = argparse.ArgumentParser(); ap.add_argument("--site", default="l3.conv2"); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--n", type=int, default=4)
a = ap.parse_args()
#            H,  C,    N,   k, stride, pad
S = {"l1.conv2": (56, 64, 64, 3, 1, 1), "l2.conv2": (28, 128, 128, 3, 1, 1), "l3.conv2": (14, 256, 256, 3, 1, 1),
     "l4.conv2": (7, 512, 512, 3, 1, 1), "l3.down": (28, 512, 1024, 1, 2, 0), "l3.conv2s2": (28, 256, 256, 3, 2, 1)}
H, C, N, k, s, p = S[a.site]
B = a.batch
x = torch.randn(B * H * H, C, device="cuda").bfloat16()
w = (torch.randn(N, k * k * C, device="cuda") * (k * k * C) ** -0.5).bfloat16()
b = torch.randn(N, device="cuda")
for _ in range(a.n): ops.conv2d_nhwc(x, w, b, B, H, H, k, k, s, p, act=ops.ACT_RELU)
torch.cuda.synchronize()
