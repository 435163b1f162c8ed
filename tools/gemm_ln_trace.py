"""clock64 pipeline trace of CTA 0 of linear + residual + LayerNorm (debug hook mvlt_debug_gemm_ln_trace).  argv: M K [warm]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import _lib, ops
lib = _lib.ensure_init()
M, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8384, 768)
warm = len(sys.argv) > 3
N = 768
def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()
a = rnd(M, K, seed=1).bfloat16(); res = rnd(M, N, seed=2)
w, bias = rnd(N, K, seed=3, scale=K ** -0.5).bfloat16(), rnd(N, seed=4, scale=0.1)
g, b = 1 + rnd(N, seed=5, scale=0.1), rnd(N, seed=6, scale=0.1)
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.linear_residual_layernorm(a, w, bias, res, g, b, 1e-12, out=res)
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
fn = lib.mvlt_debug_gemm_ln_trace; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
fn(buf.data_ptr())
if not warm: big.zero_()                                     # flush L2
ops.linear_residual_layernorm(a, w, bias, res, g, b, 1e-12, out=res)
torch.cuda.synchronize()
fn(None)
t = buf.cpu().tolist(); t0 = t[0]
rel = lambda i: (t[i] - t0) if t[i] else -1
print(f"M={M} K={K} {'warm L2' if warm else 'cold L2'} (cycles from kernel start, CTA 0; last tile of the CTA)")
print(f"setup done {rel(1)} | mma: tmem free {rel(8)} first stage full {rel(9)} all issued {rel(10)}")
print(f"epilogue warp 0: acc full {rel(16)} residual landed {rel(17)} pass1 done {rel(18)} mean {rel(19)} pass2 done {rel(20)} rstd {rel(21)} "
      f"pass3 stores issued {rel(22)} stores drained {rel(23)}")
