"""clock64 pipeline trace of CTA 0 of the fused Swin block-tail kernel (debug hook mvlt_debug_tail_trace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import _lib, ops
lib = _lib.ensure_init()
M, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (12544, 384)
def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()
x = rnd(M, C, seed=1); o = rnd(M, C, seed=2).bfloat16()
wp, bp = rnd(C, C, seed=3, scale=C ** -0.5).bfloat16(), rnd(C, seed=4, scale=0.1)
g, b = 1 + rnd(C, seed=5, scale=0.1), rnd(C, seed=6, scale=0.1)
w1, b1 = rnd(4 * C, C, seed=7, scale=C ** -0.5).bfloat16(), rnd(4 * C, seed=8, scale=0.1)
w2, b2 = rnd(C, 4 * C, seed=9, scale=(4 * C) ** -0.5).bfloat16(), rnd(C, seed=10, scale=0.1)
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.swin_block_tail(x, o, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
fn = lib.mvlt_debug_tail_trace; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
fn(buf.data_ptr())
big.zero_()                                     # flush L2
ops.swin_block_tail(x, o, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
torch.cuda.synchronize()
fn(None)
t = buf.cpu().tolist()
t0 = t[0]
rel = lambda i: (t[i] - t0) if t[i] else -1
print(f"M={M} C={C}  (cycles from the MMA warp's start)")
print(f"mma: a0_full {rel(1)}  mma0 issued {rel(2)}  a1_full {rel(3)}")
print(f"cw0: start {rel(4)} acc0_full {rel(5)} pass1 {rel(6)} pass2 {rel(7)} a1 written {rel(8)} acc2_full {rel(9)} end {rel(10)}")
n = 4 * C // 128
for j in range(n):
    print(f"  chunk {j:2d}: mma loop top {rel(16 + 4 * j)}  mma1(j+1) issued {rel(16 + 4 * j + 1)}  mma2(j) issued {rel(16 + 4 * j + 2)}   | gelu: acc1_full {rel(80 + 4 * j)}  a2 written {rel(80 + 4 * j + 1)}")
