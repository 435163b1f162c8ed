"""Device time of the one-BertLayer forward-with-saved-activations and backward (training.py) at the bench shape, per C-ABI call."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import _lib, synth, training
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
B, L = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 80
S, D = 51 + L, 768
lib = _lib.ensure_init()
model = M.MVLBertForVQA(C.offline_config("vqa", max_length=L)).eval()
sd = synth.load_synth(model, 0, "stress")
w = training.pack_layer(sd, "MVLBert.encoder.layer.0.")
g = torch.Generator().manual_seed(0)
h = torch.randn(B * S, D, generator=g).cuda(); dout = (torch.randn(B * S, D, generator=g) * 0.1).cuda()
kmask = torch.zeros(B, S, device="cuda")
for _ in range(3):
    out, saved = training.bert_layer_forward(w, h, kmask, B, S); dh, grads = training.bert_layer_backward(w, saved, dout)
torch.cuda.synchronize()
def timed(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r
tf, (out, saved) = timed(lambda: training.bert_layer_forward(w, h, kmask, B, S))
tb, _ = timed(lambda: training.bert_layer_backward(w, saved, dout))
# per entry point: wrap the C ABI with events (eager, includes launch gaps)
names = [n for n in _lib.PROTOTYPES if n not in ("mvlt_init", "mvlt_abi_version", "mvlt_linear_ln_resident_tiles")]
orig = {n: getattr(lib, n) for n in names}; ev = []
def wrap(n):
    f = orig[n]
    def gfn(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a); e1.record(); ev.append((n, e0, e1)); return r
    return gfn
for n in names: setattr(lib, n, wrap(n))
training.bert_layer_backward(w, saved, dout); torch.cuda.synchronize()
for n in names: setattr(lib, n, orig[n])
agg = collections.OrderedDict()
for n, e0, e1 in ev:
    d = agg.setdefault(n, [0, 0.0]); d[0] += 1; d[1] += e0.elapsed_time(e1)
fl_fwd = 2.0 * B * S * D * (3 * D + D + 4 * D + 4 * D) + 4.0 * B * 12 * S * S * 64
print(f"batch {B}, S = {S}: forward (unfused, saves activations, fc1 twice) {tf * 1e3:.0f} us; backward {tb * 1e3:.0f} us "
      f"({2 * fl_fwd / tb / 1e9:.0f} TFLOP/s on 2x the forward FLOPs); inference forward of the same layer (fused kernels): ~170 us")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"   {t * 1e3:8.1f} us  x{c:2d}  {n}")
