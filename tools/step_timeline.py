"""Per-kernel timeline of one forward INSIDE the replayed CUDA graph: an external timing event is recorded in front of
every C-ABI call during capture; after a replay the gaps between consecutive events are the in-graph durations (kernel +
its launch gap).  The event nodes perturb the graph a little (no programmatic overlap across them): compare the sum with
the unperturbed replay time printed last."""
import argparse, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import _lib, synth
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--len", type=int, default=80)
ap.add_argument("--reps", type=int, default=5); ap.add_argument("--detail", action="store_true")
ap.add_argument("--conv", default="swintransformer")
a = ap.parse_args()
lib = _lib.ensure_init()
torch.manual_seed(0)
model = M.MVLBertForVQA(C.offline_config("vqa", max_length=a.len, conv=a.conv)).eval().to("cuda")
model.set_precision("bf16")
x = synth.synth_images(a.batch, 1, 0.02).cuda(); ids = synth.synth_token_ids(a.batch, a.len, 1).cuda()
names = [n for n in _lib.PROTOTYPES if n not in ("mvlt_init", "mvlt_abi_version")]
orig = {n: getattr(lib, n) for n in names}
events, labels = [], []
def label(n, args):
    if n == "mvlt_gemm_bf16_tc": return f"gemm {args[10]}x{args[11]}x{args[12]} act{args[13]} out{args[14]} res{int(args[7] is not None)}"
    if n == "mvlt_conv2d_nhwc_bf16_tc":
        return f"conv H{args[2]} C{args[4]} N{args[12]} k{args[13]} s{args[15]} act{args[17]} res{int(args[10] is not None)}"
    if n == "mvlt_layernorm_rows": return f"layernorm rows{args[8]} C{args[9]}"
    if n == "mvlt_window_attention": return f"window_attn H{args[5]} C{args[7]} shift{args[10]}"
    if n == "mvlt_swin_mlp_fused": return f"swin_mlp_fused M{args[9]} C{args[10]}"
    if n == "mvlt_swin_ln_qkv": return f"swin_ln_qkv M{args[8] * args[9] * args[10]} C{args[11]} N{args[12]}"
    if n == "mvlt_swin_block_tail": return f"swin_block_tail M{args[12]} C{args[13]}"
    if n == "mvlt_window_attention_tc": return f"window_attn_tc H{args[4]} C{args[6]} shift{args[9]}"
    if n == "mvlt_layernorm_rows_winmajor": return f"layernorm_winmajor rows{args[5] * args[6] * args[7]} C{args[8]}"
    if n == "mvlt_linear_residual_layernorm": return f"linear_residual_layernorm {args[14]}x{args[15]}x{args[16]}"
    if n == "mvlt_joint_attention_tc": return f"joint_attention_tc S{args[4]}"
    return n.replace("mvlt_", "")
def wrap(n):
    f = orig[n]
    def g(*args):
        e = torch.cuda.Event(enable_timing=True, external=True); e.record(); events.append(e); labels.append(label(n, args))
        return f(*args)
    return g
st = torch.cuda.Stream()
with torch.cuda.stream(st), torch.no_grad():
    for _ in range(2): model(x, ids, None)
    st.synchronize()
    g0 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g0, stream=st): model(x, ids, None)
    for n in names: setattr(lib, n, wrap(n))
    g1 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1, stream=st):
        model(x, ids, None)
        e = torch.cuda.Event(enable_timing=True, external=True); e.record(); events.append(e)
    for n in names: setattr(lib, n, orig[n])
    acc = [0.0] * len(labels); tot = 0.0
    for _ in range(2): g1.replay()
    st.synchronize()
    for _ in range(a.reps):
        g1.replay(); st.synchronize()
        for i in range(len(labels)): acc[i] += events[i].elapsed_time(events[i + 1])
        tot += events[0].elapsed_time(events[-1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): g0.replay()
    e0.record(st)
    for _ in range(10): g0.replay()
    e1.record(st); st.synchronize()
plain = e0.elapsed_time(e1) / 10
by = collections.OrderedDict()
for l, t in zip(labels, acc):
    d = by.setdefault(l, [0, 0.0]); d[0] += 1; d[1] += t / a.reps
print(f"{len(labels)} launches; instrumented replay {tot / a.reps:.3f} ms; plain graph replay {plain:.3f} ms")
for l, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{t*1e3:9.1f} us {100*t/(tot/a.reps):5.1f}%  x{n:3d}  {t*1e3/n:7.1f} us each   {l}")
if a.detail:
    for l, t in zip(labels, acc): print(f"   {t/a.reps*1e3:8.1f}  {l}")
