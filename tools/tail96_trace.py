"""clock64 pipeline trace of CTA 0 of the persistent stage-0 block-tail kernel (debug hook mvlt_debug_tail96_trace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import _lib, ops
lib = _lib.ensure_init()
M, C = (int(sys.argv[1]) if len(sys.argv) > 1 else 200704), 96
def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()
x = rnd(M, C, seed=1); o = rnd(M, C, seed=2).bfloat16()
wp, bp = rnd(C, C, seed=3, scale=C ** -0.5).bfloat16(), rnd(C, seed=4, scale=0.1)
g, b = 1 + rnd(C, seed=5, scale=0.1), rnd(C, seed=6, scale=0.1)
w1, b1 = rnd(4 * C, C, seed=7, scale=C ** -0.5).bfloat16(), rnd(4 * C, seed=8, scale=0.1)
w2, b2 = rnd(C, 4 * C, seed=9, scale=(4 * C) ** -0.5).bfloat16(), rnd(C, seed=10, scale=0.1)
for _ in range(3): ops.swin_block_tail(x, o, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
fn = lib.mvlt_debug_tail96_trace; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
fn(buf.data_ptr())
ops.swin_block_tail(x, o, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
torch.cuda.synchronize(); fn(None)
t = buf.cpu().tolist(); t0 = t[0]
rel = lambda i: (t[i] - t0) if t[i] else -1
print(f"M={M}: mma warp start 0 .. end {rel(1)} cycles (CTA 0, {(M + 255) // 256} tiles over the grid)")
for i in range(4):
    print(f" tile {i}: LN group: slot free {rel(32 + 4 * i)}  x in TMEM {rel(33 + 4 * i)}  proj seen {rel(34 + 4 * i)}  A1 written {rel(35 + 4 * i)} | "
          f"store: acc2 full {rel(96 + 2 * i)} done {rel(97 + 2 * i)}")
for g_ in range(12):
    print(f"  chunk {g_:2d}: fc2 issued {rel(16 + g_)} | gelu: acc1 full {rel(64 + 2 * g_)}  a2 written {rel(65 + 2 * g_)}")
