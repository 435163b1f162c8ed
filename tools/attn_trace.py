"""clock64 pipeline trace of CTA 0 of the tcgen05 joint-attention kernel (debug hook mvlt_debug_attn_trace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import _lib, ops
lib = _lib.ensure_init()
B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 131)
g = torch.Generator(device="cpu").manual_seed(0)
qkv = torch.randn(B * S, 2304, generator=g).cuda().bfloat16()
kmask = torch.zeros(B, S, device="cuda")
for _ in range(3):
    ops.joint_attention(qkv, kmask, B, S, 12, False, 50, impl="tc")
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
fn = lib.mvlt_debug_attn_trace; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
fn(buf.data_ptr())
ops.joint_attention(qkv, kmask, B, S, 12, False, 50, impl="tc")
torch.cuda.synchronize()
fn(None)
t = buf.cpu().tolist()
t0 = t[0]
print(f"kernel body {t[1] - t0} cycles")
names = ["qk_tma_issue", "v_tma_issue", "S_mma_issue", "PV_mma_issue", "mask_filled", "mask_bar", "s_full_seen", "pass1_done", "p_written", "o_full_seen", "epi_done", "S_last_not_ready", "PV_last_not_ready"]
for n in range(8):
    row = t[16 + n * 16: 16 + n * 16 + 13]
    if not any(row):
        continue
    print(f"tile {n} (slot {n % 2}): " + "  ".join(f"{nm} {v - t0 if v else -1}" for nm, v in zip(names, row)))
