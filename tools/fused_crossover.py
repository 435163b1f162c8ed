"""Fused kernel vs the unfused chain as a function of the batch (rows): where does each one-wave design stop paying?
In-graph, 20 launches back to back, four rotated residual buffers.  python tools/fused_crossover.py [batches...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from medical_vision_langauge_transformer_b200 import ops

def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()
def graph_time(fn, n=20):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); g.replay(); e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

batches = [int(a) for a in sys.argv[1:]] or [8, 16, 24, 32, 40, 48, 64]
for B in batches:
    # BERT post-LN sites: rows = B * 131
    M, N = B * 131, 768
    line = f"batch {B:3d}: "
    for K in (768, 3072):
        a = rnd(M, K, seed=1).bfloat16(); w, bias = rnd(N, K, seed=3, scale=K ** -0.5).bfloat16(), rnd(N, seed=4, scale=0.1)
        g, b = 1 + rnd(N, seed=5, scale=0.1), rnd(N, seed=6, scale=0.1)
        xs = [rnd(M, N, seed=7 + i) for i in range(4)]; k = [0]
        def fused():
            k[0] = (k[0] + 1) % 4; ops.linear_residual_layernorm(a, w, bias, xs[k[0]], g, b, 1e-12, out=xs[k[0]])
        def chain():
            k[0] = (k[0] + 1) % 4; xc = xs[k[0]]
            ops.linear(a, w, bias, residual=xc, out=xc); ops.layernorm(xc, g, b, 1e-12, torch.float32, out=xc, bf16_copy=True)
        line += f"gemm_ln K={K}: {graph_time(fused):5.1f} vs {graph_time(chain):5.1f} us | "
    # Swin stage 2 (C = 384, rows = B * 196) and stage 1 (C = 192, rows = B * 784)
    for C, H in ((384, 14), (192, 28)):
        M = B * H * H
        o = rnd(M, C, seed=2).bfloat16()
        wp, bp = rnd(C, C, seed=3, scale=C ** -0.5).bfloat16(), rnd(C, seed=4, scale=0.1)
        g, b = 1 + rnd(C, seed=5, scale=0.1), rnd(C, seed=6, scale=0.1)
        w1, b1 = rnd(4 * C, C, seed=7, scale=C ** -0.5).bfloat16(), rnd(4 * C, seed=8, scale=0.1)
        w2, b2 = rnd(C, 4 * C, seed=9, scale=(4 * C) ** -0.5).bfloat16(), rnd(C, seed=10, scale=0.1)
        wq, bq = rnd(3 * C, C, seed=11, scale=C ** -0.5).bfloat16(), rnd(3 * C, seed=12, scale=0.1)
        xs = [rnd(M, C, seed=20 + i) for i in range(4)]; k = [0]
        def tail_f():
            k[0] = (k[0] + 1) % 4; ops.swin_block_tail(xs[k[0]], o, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
        def tail_c():
            k[0] = (k[0] + 1) % 4; xc = xs[k[0]]
            ops.linear(o, wp, bp, residual=xc, out=xc); an = ops.layernorm(xc, g, b, 1e-5, torch.bfloat16)
            hn = ops.linear(an, w1, b1, act=ops.ACT_GELU); ops.linear(hn, w2, b2, residual=xc, out=xc)
        def lq_f():
            k[0] = (k[0] + 1) % 4; ops.swin_ln_qkv(xs[k[0]], g, b, 1e-5, wq, bq, B, H, H, 7, 3)
        def lq_c():
            k[0] = (k[0] + 1) % 4; ops.linear(ops.layernorm_winmajor(xs[k[0]], g, b, 1e-5, B, H, H, 7, 3), wq, bq)
        line += f"C={C} tail: {graph_time(tail_f):5.1f} vs {graph_time(tail_c):5.1f}  ln_qkv: {graph_time(lq_f):5.1f} vs {graph_time(lq_c):5.1f} | "
    print(line, flush=True)
