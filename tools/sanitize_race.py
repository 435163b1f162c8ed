"""bf16 forwards (Swin + ResNet-50 retrieval, batch 1) for compute-sanitizer --tool racecheck / synccheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import synth
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
x, ids = synth.synth_images(1, 1, 1.0).cuda(), synth.synth_token_ids(1, 80, 1).cuda()
with torch.no_grad():
    for conv in ("swintransformer", "resnet50"):
        m = M.MVLBertForRetrieval(C.offline_config("retrieval", conv=conv, max_length=80)).eval()
        synth.load_synth(m, 0, "stress"); m = m.cuda().set_precision("bf16")
        print(conv, m(x, ids).flatten().tolist())
torch.cuda.synchronize()
