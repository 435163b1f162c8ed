"""Joint / window attention time versus batch (waves of CTAs): us per launch and per 64 samples, events around 20 launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
S = 131
for B in (16, 32, 64, 128, 256, 512):
    qkv = torch.randn(B * S, 2304, device="cuda").bfloat16(); km = torch.zeros(B, S, device="cuda"); out = torch.empty(B * S, 768, device="cuda", dtype=torch.bfloat16)
    us = timed(lambda: ops.joint_attention(qkv, km, B, S, 12, False, 50, out=out))
    print(f"joint  B={B:4d}  {us:8.1f} us  -> {us * 64 / B:6.1f} us per 64 samples   ({B * 12} CTAs)")
for B in (16, 64, 256):
    H, C, heads = 14, 384, 12
    qkv = torch.randn(B * H * H, 3 * C, device="cuda").bfloat16()
    frag = ops.window_bias_fragments(torch.randn(heads, 64, 64, device="cuda"), 3, 32 ** -0.5)
    out = torch.empty(B * H * H, C, device="cuda", dtype=torch.bfloat16)
    us = timed(lambda: ops.window_attention(qkv, frag, B, H, H, C, heads, 7, 3, 32 ** -0.5, out=out))
    print(f"window s2 B={B:4d}  {us:8.1f} us  -> {us * 64 / B:6.1f} us per 64 samples")
