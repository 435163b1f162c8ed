"""Pipeline trace of CTA 0 of the fused Swin MLP kernel (clock64 stamps written through the mvlt_debug_mlp_trace hook)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops, _lib
lib = _lib.ensure_init()
lib.mvlt_debug_mlp_trace.argtypes = [ctypes.c_void_p]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 384
M = int(sys.argv[2]) if len(sys.argv) > 2 else 12544
x = torch.randn(M, C, device="cuda"); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
w1 = (torch.randn(4 * C, C, device="cuda") * C ** -0.5).bfloat16(); b1 = torch.randn(4 * C, device="cuda") * 0.1
w2 = (torch.randn(C, 4 * C, device="cuda") * (4 * C) ** -0.5).bfloat16() * 0.1; b2 = torch.randn(C, device="cuda") * 0.1
for _ in range(3): ops.swin_mlp(x, g, b, 1e-5, w1, b1, w2, b2)
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib.mvlt_debug_mlp_trace(tr.data_ptr())
ops.swin_mlp(x, g, b, 1e-5, w1, b1, w2, b2)
torch.cuda.synchronize()
lib.mvlt_debug_mlp_trace(None)
t = tr.cpu().tolist()
t0 = min(v for v in t[:8] if v)
rel = lambda i: t[i] - t0 if t[i] else None
print("C", C, "M", M)
print("mma: start", rel(0), "a1_full", rel(1), "| epi: start", rel(2), "pdl", rel(3), "ln done", rel(4), "acc2_full", rel(5), "end", rel(6))
n = 4 * C // 64
print(" j | mma: m1(j+1) issued, a2_full(j) seen, m2(j) issued | epi0: acc1_full, gelu done, a2_empty, arrived")
for j in range(n):
    print(f"{j:2d} | {rel(16+4*j)} {rel(17+4*j)} {rel(18+4*j)} | {rel(256+4*j)} {rel(257+4*j)} {rel(258+4*j)} {rel(259+4*j)}")
