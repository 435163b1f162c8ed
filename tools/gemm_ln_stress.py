"""Row-placement independence of linear + residual + LayerNorm: rows computed inside a big matrix == the same rows computed alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import ops
N = 768
for K in (768, 3072):
    g0 = torch.Generator(device="cpu").manual_seed(K)
    M = 268288
    a = (torch.randn(M, K, generator=g0)).cuda().bfloat16(); res = torch.randn(M, N, generator=g0).cuda()
    w = (torch.randn(N, K, generator=g0) * K ** -0.5).cuda().bfloat16(); bias = torch.randn(N, generator=g0).cuda() * 0.1
    g = 1 + 0.1 * torch.randn(N, generator=g0).cuda(); b = 0.1 * torch.randn(N, generator=g0).cuda()
    full, fs = ops.linear_residual_layernorm(a, w, bias, res, g, b, 1e-12)
    for rep in range(3):
        again, _ = ops.linear_residual_layernorm(a, w, bias, res, g, b, 1e-12)
        d = (again != full).any(1).nonzero().flatten()
        print(f"K={K} rerun {rep}: rows differing {d.numel()} {d[:8].tolist()}")
    for lo, n in ((0, 67072), (131 * 700, 67072), (200000, 8384), (1000, 300)):
        part, _ = ops.linear_residual_layernorm(a[lo:lo + n], w, bias, res[lo:lo + n], g, b, 1e-12)
        d = (part != full[lo:lo + n]).any(1).nonzero().flatten()
        md = (part - full[lo:lo + n]).abs().max().item()
        print(f"K={K} rows [{lo}, +{n}) alone vs inside: rows differing {d.numel()} max|d| {md:.2e} first {d[:8].tolist()} tiles {sorted(set((d // 256).tolist()))[:10]}")
