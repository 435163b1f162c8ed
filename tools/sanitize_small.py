"""Small forwards of every backbone + heads for compute-sanitizer (memcheck): batch 2, bf16 and fp32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from medical_vision_langauge_transformer_b200 import synth, retrieval
from medical_vision_langauge_transformer_b200.modules import config as C, model as M
x, ids = synth.synth_images(2, 1, 1.0).cuda(), synth.synth_token_ids(2, 80, 1).cuda()
with torch.no_grad():
    for conv in ("swintransformer", "resnet50", "linear", "vit"):
        m = M.MVLBertForRetrieval(C.offline_config("retrieval", conv=conv, max_length=80)).eval()
        synth.load_synth(m, 0, "stress"); m = m.cuda()
        for prec in ("bf16", "fp32"):
            p = m.set_precision(prec)(x, ids)
            torch.cuda.synchronize(); print(conv, prec, p.flatten().tolist())
        if conv == "swintransformer":
            s, met = retrieval.rank_task(m.set_precision("bf16"), x.cpu(), ids.cpu(), torch.eye(2), pair_batch=4); print("rank", met)
            x5 = torch.cat([x, x]).view(2, 2, 3, 224, 224); print("two-view", m(x5, ids).flatten().tolist())
    pm = M.MVLBertForPretraining(C.offline_config("pretrain", max_length=80, ITM_task=True)).eval(); synth.load_synth(pm, 0, "stress"); pm = pm.cuda()
    masked, labels = synth.synth_mlm_labels(ids.cpu(), 3)
    print("pretrain", pm(x, masked.cuda(), labels.cuda(), torch.tensor([1, 0]).cuda()).item())
    cm = M.MVLBertForImageCaption(C.offline_config("caption", max_length=4)).eval(); synth.load_synth(cm, 0, "stress"); cm = cm.cuda()
    print("greedy", cm(x, None, 1, "unilm")[0].tolist())
    # first slice of the training step: one BertLayer forward-with-saved-activations + backward (csrc/backward.cu)
    from medical_vision_langauge_transformer_b200 import training
with torch.no_grad():
    vm = M.MVLBertForVQA(C.offline_config("vqa", max_length=80)).eval()
    w = training.pack_layer(synth.load_synth(vm, 0, "stress"), "MVLBert.encoder.layer.0.")
    h = torch.randn(2 * 131, 768, device="cuda")
    for s2s in (False, True):
        out, saved = training.bert_layer_forward(w, h, None, 2, 131, 12, s2s, 50)
        dh, grads = training.bert_layer_backward(w, saved, torch.randn_like(h) * 0.1)
        print("backward", s2s, dh.abs().max().item(), len(grads))
torch.cuda.synchronize()
