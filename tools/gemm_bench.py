"""GEMM micro-benchmark over the real call-site shapes of the batch-64 step (run on the GPU box)."""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import ops
ap = argparse.ArgumentParser(); ap.add_argument("--sweep", action="store_true"); ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
# (name, M, N, K, act, out_dtype, res_dtype)
F32, BF = torch.float32, torch.bfloat16
SHAPES = [
    ("s0.qkv", 200704, 288, 96, 0, BF, None), ("s0.proj", 200704, 96, 96, 0, F32, F32), ("s0.fc1", 200704, 384, 96, 1, BF, None), ("s0.fc2", 200704, 96, 384, 0, F32, F32),
    ("s1.qkv", 50176, 576, 192, 0, BF, None), ("s1.proj", 50176, 192, 192, 0, F32, F32), ("s1.fc1", 50176, 768, 192, 1, BF, None), ("s1.fc2", 50176, 192, 768, 0, F32, F32),
    ("s2.qkv", 12544, 1152, 384, 0, BF, None), ("s2.proj", 12544, 384, 384, 0, F32, F32), ("s2.fc1", 12544, 1536, 384, 1, BF, None), ("s2.fc2", 12544, 384, 1536, 0, F32, F32),
    ("s3.qkv", 3136, 2304, 768, 0, BF, None), ("s3.proj", 3136, 768, 768, 0, F32, F32), ("s3.fc1", 3136, 3072, 768, 1, BF, None), ("s3.fc2", 3136, 768, 3072, 0, F32, F32),
    ("bert.qkv", 8384, 2304, 768, 0, BF, None), ("bert.ao", 8384, 768, 768, 0, F32, F32), ("bert.fi", 8384, 3072, 768, 1, BF, None), ("bert.fo", 8384, 768, 3072, 0, F32, F32),
]
COUNT = {"s0": 2, "s1": 2, "s2": 18, "s3": 2, "bert": 12}
def run(M, N, K, act, od, rd, bn):
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, device="cuda"); r = None if rd is None else torch.randn(M, N, device="cuda").to(rd)
    out = torch.empty(M, N, device="cuda", dtype=od)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): ops.linear(x, w, b, act=act, residual=r, out=out, block_n=bn)
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(x, w, b, act=act, residual=r, out=out, block_n=bn); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
total = 0.0
for name, M, N, K, act, od, rd in SHAPES:
    res = {}
    for bn in ([0] + ([32, 64, 96, 128, 192, 256] if a.sweep else [])):
        if bn > 0 and bn > ((N + 31) // 32) * 32: continue
        ms = run(M, N, K, act, od, rd, bn); res[bn] = ms
    best = min(res.values()); total += COUNT[name.split(".")[0]] * res[0]
    print(f"{name:9s} {M:7d}x{N:5d}x{K:5d} auto {res[0]*1e3:8.1f} us {2*M*N*K/res[0]/1e9:7.1f} TF/s | " +
          " ".join(f"bn{bn}:{2*M*N*K/ms/1e9:6.0f}" for bn, ms in res.items() if bn), flush=True)
print(f"sum over the step (auto tile): {total:.3f} ms")
