"""Do two forward steps in flight (two CUDA graphs with separate memory pools on two streams) fill each other's kernel
tails?  Prints ms/step for 1 and 2 (and 3) concurrent streams."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import synth, _lib
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
_lib.ensure_init()
torch.manual_seed(0)
B, L = 64, 80
model = M.MVLBertForVQA(C.offline_config("vqa", max_length=L)).eval().cuda().set_precision("bf16")
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 3
xs = [synth.synth_images(B, 10 + i, 0.02).cuda() for i in range(NS)]; ids = [synth.synth_token_ids(B, L, 10 + i).cuda() for i in range(NS)]
streams = [torch.cuda.Stream() for _ in range(NS)]
graphs, outs = [], []
with torch.no_grad():
    for i in range(NS):
        with torch.cuda.stream(streams[i]):
            for _ in range(2): model(xs[i], ids[i], None)
            streams[i].synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[i]):          # private pool per graph
                outs.append(model(xs[i], ids[i], None))
            graphs.append(g)
for i in range(NS):
    with torch.cuda.stream(streams[i]): graphs[i].replay()
torch.cuda.synchronize()
ref = [o[1].clone() for o in outs]
def run(n_streams, steps=40):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams[:n_streams]: s.wait_event(e0)
    for k in range(steps):
        i = k % n_streams
        with torch.cuda.stream(streams[i]): graphs[i].replay()
    for s in streams[:n_streams]: main.wait_stream(s)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for n in range(1, NS + 1):
    run(n, 10)
    print(f"{n} stream(s): {run(n):.3f} ms/step  ({B / run(n) * 1e3:.0f} pairs/s)")
for i in range(NS):
    assert torch.equal(outs[i][1], ref[i]), "concurrent replays changed the logits"
print("logits identical to the serial replays")
