"""clock64 pipeline trace of CTA 0 of the fused LayerNorm + qkv kernel (debug hook mvlt_debug_lnqkv_trace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import _lib, ops
lib = _lib.ensure_init()
B, H, C = 64, 14, 384
g0 = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(B * H * H, C, generator=g0).cuda()
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
w = (torch.randn(3 * C, C, generator=g0) * C ** -0.5).cuda().bfloat16(); bias = torch.zeros(3 * C, device="cuda")
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.swin_ln_qkv(x, g, b, 1e-5, w, bias, B, H, H, 7, 3)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
fn = lib.mvlt_debug_lnqkv_trace; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
for flush in (False, True):
    buf.zero_(); fn(buf.data_ptr())
    if flush: big.zero_()
    ops.swin_ln_qkv(x, g, b, 1e-5, w, bias, B, H, H, 7, 3)
    torch.cuda.synchronize(); fn(None)
    t = buf.cpu().tolist(); t0 = t[0]; rel = lambda i: (t[i] - t0) if t[i] else -1
    print(f"L2 {'flushed' if flush else 'warm'}: mma: a1_full {rel(1)} | cw0: after pdl {rel(2)} LN done {rel(3)} end {rel(4)}")
    for nc in range(5):
        print(f"   chunk {nc}: mma start {rel(16 + 2 * nc)} issued {rel(16 + 2 * nc + 1)} | epi: acc_full {rel(48 + 2 * nc)} drained {rel(48 + 2 * nc + 1)}")
