"""One eager forward step of the bench workload (for ncu): `--passes P` forwards of MVLBertForVQA, batch B, L=80, bf16.
Under `ncu -s <launches of the first passes>` the last pass is the profiled step."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import synth, runtime
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--max-length", type=int, default=80); ap.add_argument("--count-only", action="store_true")
ap.add_argument("--conv", default="swintransformer")
a = ap.parse_args()
torch.manual_seed(0)
model = M.MVLBertForVQA(C.offline_config("vqa", max_length=a.max_length, conv=a.conv)).eval().cuda()
x, ids = synth.synth_images(a.batch, 1, 0.02).cuda(), synth.synth_token_ids(a.batch, a.max_length, 1).cuda()
with torch.no_grad():
    if a.count_only:
        model(x, ids, None); print("launches_per_step", runtime.count_launches(lambda: model(x, ids, None)))
    else:
        for _ in range(a.passes):
            model(x, ids, None)
torch.cuda.synchronize()
