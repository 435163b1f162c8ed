"""Launch the attention kernels of one stage a few times (for ncu --set full)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops
ap = argparse.ArgumentParser(); ap.add_argument("--what", default="s2"); ap.add_argument("--n", type=int, default=4); ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
B = a.batch
if a.what == "joint":
    S = 131; qkv = torch.randn(B * S, 2304, device="cuda").bfloat16(); km = torch.zeros(B, S, device="cuda")
    for _ in range(a.n): ops.joint_attention(qkv, km, B, S, 12, False, 50)
else:
    H, C, heads = {"s0": (56, 96, 3), "s1": (28, 192, 6), "s2": (14, 384, 12), "s3": (7, 768, 24)}[a.what]
    qkv = torch.randn(B * H * H, 3 * C, device="cuda").bfloat16()
    frag = ops.window_bias_fragments(torch.randn(heads, 64, 64, device="cuda"), 3 if H > 7 else 0, 32 ** -0.5)
    for _ in range(a.n): ops.window_attention(qkv, frag, B, H, H, C, heads, 7, 3 if H > 7 else 0, 32 ** -0.5)
torch.cuda.synchronize()
