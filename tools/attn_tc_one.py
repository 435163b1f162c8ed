"""One configuration of a tcgen05 attention kernel, a few launches (for ncu): attn_tc_one.py window B H C heads shift | joint B S"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import ops
kind = sys.argv[1]
g = torch.Generator(device="cpu").manual_seed(0)
if kind == "window":
    B, H, C, heads, shift = map(int, sys.argv[2:7])
    qkv = torch.randn(B * H * H, 3 * C, generator=g).cuda().bfloat16()
    relb = torch.zeros(heads, 64, 64, device="cuda")
    table = ops.window_bias_table(relb, shift)
    for _ in range(4):
        ops.window_attention_tc(qkv, table, B, H, H, C, heads, 7, shift, 32 ** -0.5)
else:
    B, S = map(int, sys.argv[2:4])
    qkv = torch.randn(B * S, 2304, generator=g).cuda().bfloat16()
    kmask = torch.zeros(B, S, device="cuda")
    for _ in range(4):
        ops.joint_attention(qkv, kmask, B, S, 12, False, 50, impl="tc")
torch.cuda.synchronize()
