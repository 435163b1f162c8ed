"""Launch one GEMM call-site shape a few times (for ncu --set full)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops
ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="s0.qkv"); ap.add_argument("--bn", type=int, default=0); ap.add_argument("--n", type=int, default=4)
a = ap.parse_args()
F32, BF = torch.float32, torch.bfloat16
S = {"s0.qkv": (200704, 288, 96, 0, BF, None), "s0.proj": (200704, 96, 96, 0, F32, F32), "s0.fc1": (200704, 384, 96, 1, BF, None),
     "s2.qkv": (12544, 1152, 384, 0, BF, None), "s2.proj": (12544, 384, 384, 0, F32, F32), "s2.fc1": (12544, 1536, 384, 1, BF, None), "s2.fc2": (12544, 384, 1536, 0, F32, F32),
     "bert.qkv": (8384, 2304, 768, 0, BF, None), "bert.ao": (8384, 768, 768, 0, F32, F32), "bert.fi": (8384, 3072, 768, 1, BF, None), "bert.fo": (8384, 768, 3072, 0, F32, F32)}
M, N, K, act, od, rd = S[a.shape]
x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
b = torch.randn(N, device="cuda"); r = None if rd is None else torch.randn(M, N, device="cuda").to(rd)
out = r if (rd == od and r is not None) else torch.empty(M, N, device="cuda", dtype=od)
for _ in range(a.n): ops.linear(x, w, b, act=act, residual=r, out=out, block_n=a.bn)
torch.cuda.synchronize()
