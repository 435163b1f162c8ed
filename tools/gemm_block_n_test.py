import os, sys
sys.path.insert(0, "/root/repo")
import torch
from medical_vision_langauge_transformer_b200 import ops
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
for (M, N, K, act) in [(8384, 2304, 768, 0), (8384, 3072, 768, 1), (12544, 1152, 384, 0), (3136, 2304, 768, 0), (3136, 3072, 768, 1), (67072, 2304, 768, 0)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
    outs = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(4)]; k = [0]
    ref = ops.linear(a, w, b, act=act)
    line = f"{M}x{N}x{K} act{act}: "
    for bn in (0, 256, 224, 192, 160, 128):
        def f():
            k[0] = (k[0] + 1) % 4; ops.linear(a, w, b, act=act, out=outs[k[0]], block_n=bn)
        t = graph_time(f)
        ok = torch.equal(outs[k[0]], ref)
        line += f"bn{bn}: {t:5.1f} us{'' if ok else ' (!=)'}  "
    print(line, flush=True)
