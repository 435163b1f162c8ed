"""Per-tap error growth of the bf16 path vs the reference golden probes (diagnostic; run on the GPU box)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from medical_vision_langauge_transformer_b200 import synth
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
from oracle.make_golden import probe_indices
for case in ("retrieval_stress", "retrieval_config1"):
    g = torch.load(os.path.join(ROOT, "tests/golden", case + ".pt"))
    model = M.MVLBertForRetrieval(C.offline_config("retrieval", max_length=g["L"])).eval()
    synth.load_synth(model, g["weight_seed"], g["flavour"]); model.cuda()
    x = synth.synth_images(g["B"], g["data_seed"], g["img_scale"]).cuda(); ids = synth.synth_token_ids(g["B"], g["L"], g["data_seed"]).cuda()
    for prec in ("fp32", "bf16"):
        model.set_precision(prec); taps = {}
        model.conv.conv[0].taps = model.MVLBert.taps = taps
        with torch.no_grad(): logits = model(x, ids, image_text_label=1)
        line = []
        for name, gt in g["taps"].items():
            t = taps[name].float().cpu().contiguous(); v = t.flatten()[probe_indices(t.numel(), name)]
            d = (v - gt["values"])
            line.append(f"{name}: max {d.abs().max().item()/gt['absmax']:.1e} rms {d.pow(2).mean().sqrt().item()/gt['std']:.1e}")
        print(case, prec, "logits err", ((logits.cpu()-g["logits"]).abs().max()/g["logits"].abs().max()).item())
        print("   " + "\n   ".join(line))
