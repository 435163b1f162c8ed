"""Generate tests/golden/*.pt by running the REAL reference (imported via oracle/ref_shims.py).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The fixtures are small: final outputs in full, intermediate activations as 4096-element seeded probes.
Weights/inputs are NOT stored — they are regenerated from `synth` by name+seed on either side.
"""
from __future__ import annotations

import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from medical_vision_langauge_transformer_b200 import synth  # noqa: E402
from oracle.ref_shims import build_reference_model  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
PROBE_N = 4096


def probe_indices(numel: int, name: str) -> torch.Tensor:
    g = synth._gen("probe:" + name, 0)
    return torch.randint(0, numel, (min(PROBE_N, numel),), generator=g)


def probe(t: torch.Tensor, name: str) -> dict:
    t = t.detach().float().contiguous()
    idx = probe_indices(t.numel(), name)
    return {"shape": tuple(t.shape), "values": t.flatten()[idx].clone(), "mean": t.mean().item(),
            "std": t.std().item(), "absmax": t.abs().max().item()}


def attach_taps(model, taps: dict):
    """Forward hooks on the reference modules; names match oracle.mvlt_oracle taps."""
    hs = []
    swin = model.conv.conv[0]
    if hasattr(swin, "linear_patch"):     # linear-patch stem (vfe.py:47-60): its [B,768,14,14] output
        hs.append(swin.register_forward_hook(lambda m, i, o: taps.__setitem__("linear_patch", o)))
    elif hasattr(swin, "class_token"):    # ViT trunk (vfe.py:66-107): residual stream after the first and the last block
        for li in (0, 11):
            hs.append(swin.encoder.layers[li].register_forward_hook(lambda m, i, o, k=f"vit{li}": taps.__setitem__(k, o)))
    elif hasattr(swin, "layer1"):           # ResNet trunk (vfe.py:7-44): stem output after the max-pool, then each stage
        hs.append(swin.maxpool.register_forward_hook(lambda m, i, o: taps.__setitem__("stem", o)))
        for li in (1, 2, 3, 4):
            hs.append(getattr(swin, f"layer{li}").register_forward_hook(
                lambda m, i, o, k=f"layer{li}": taps.__setitem__(k, o)))
    else:
        hs.append(swin.patch_embed.register_forward_hook(lambda m, i, o: taps.__setitem__("patch_embed", o)))
        for s, layer in enumerate(swin.layers):
            for b in (0, 1):
                hs.append(layer.blocks[b].register_forward_hook(
                    lambda m, i, o, k=f"s{s}b{b}": taps.__setitem__(k, o)))
            hs.append(layer.register_forward_hook(lambda m, i, o, k=f"stage{s}": taps.__setitem__(k, o)))
    hs.append(model.conv.register_forward_hook(lambda m, i, o: taps.__setitem__("image_feature", o)))
    enc = model.MVLBert.encoder
    hs.append(enc.register_forward_pre_hook(
        lambda m, a, kw: taps.__setitem__("embedding", kw.get("hidden_states", a[0] if a else None)), with_kwargs=True))
    for l in (0, 11):
        hs.append(enc.layer[l].register_forward_hook(
            lambda m, i, o, k=f"bert{l}": taps.__setitem__(k, o[0] if isinstance(o, tuple) else o)))
    hs.append(model.MVLBert.pooler.register_forward_hook(lambda m, i, o: taps.__setitem__("pooled", o)))
    return hs


def case_retrieval(flavour, img_scale, B=2, L=80, conv="swintransformer"):
    m = build_reference_model("retrieval", max_length=L, conv=conv)
    synth.load_synth(m, 0, flavour)
    x, ids = synth.synth_images(B, 1, img_scale), synth.synth_token_ids(B, L, 1)
    taps = {}
    hs = attach_taps(m, taps)
    with torch.no_grad():
        prob = m(x, ids)
        logits = m(x, ids, image_text_label=torch.zeros(B, dtype=torch.long))
    [h.remove() for h in hs]
    return {"task": "retrieval", "flavour": flavour, "img_scale": img_scale, "B": B, "L": L, "weight_seed": 0, "conv": conv,
            "data_seed": 1, "prob": prob, "logits": logits, "taps": {k: probe(v, k) for k, v in taps.items()}}


def case_vqa(B=3, L=23):
    m = build_reference_model("vqa", max_length=L)
    synth.load_synth(m, 0, "stress")
    x, ids = synth.synth_images(B, 2, 1.0), synth.synth_token_ids(B, L, 2, min_len=4)
    taps = {}
    hs = attach_taps(m, taps)
    with torch.no_grad():
        prob, logits = m(x, ids, None)
    [h.remove() for h in hs]
    return {"task": "vqa", "flavour": "stress", "img_scale": 1.0, "B": B, "L": L, "weight_seed": 0, "data_seed": 2,
            "min_len": 4, "prob": prob, "logits": logits, "taps": {k: probe(v, k) for k, v in taps.items()}}


def case_pretrain(B=2, L=80):
    m = build_reference_model("pretrain", max_length=L, itm=True)
    synth.load_synth(m, 0, "stress")
    x, ids = synth.synth_images(B, 3, 1.0), synth.synth_token_ids(B, L, 3)
    masked, labels = synth.synth_mlm_labels(ids, 3)
    itm = torch.tensor([1, 0])[:B]
    out = {"task": "pretrain", "flavour": "stress", "img_scale": 1.0, "B": B, "L": L, "weight_seed": 0,
           "data_seed": 3, "itm_labels": itm}
    for seed in range(64):                      # find one seed per branch of model.py:390-394
        random.seed(seed)
        branch = "seq2seq" if random.random() < 0.5 else "bidir"
        if branch in out:
            continue
        random.seed(seed)
        with torch.no_grad():
            loss = m(x, masked, labels, itm)
        out[branch] = {"py_seed": seed, "loss": loss}
        if "seq2seq" in out and "bidir" in out:
            break
    return out


def case_caption(B=2, L=40):
    """MVLBertForImageCaption teacher-forced pass (num_beams=0, model.py:518-550), both learning strategies; logits are
    [B, 30522, L] — stored as probes."""
    m = build_reference_model("caption", max_length=L)
    synth.load_synth(m, 0, "stress")
    x, ids = synth.synth_images(B, 5, 1.0), synth.synth_token_ids(B, L, 5)
    out = {"task": "caption", "flavour": "stress", "img_scale": 1.0, "B": B, "L": L, "weight_seed": 0, "data_seed": 5}
    with torch.no_grad():
        for strategy in ("unilm", "normal"):
            out[strategy] = probe(m(x, ids, 0, strategy), "caption_" + strategy)
    return out


def case_rank(N=6, L=80):
    """run_retrieval.py test split: row-major N*N pair enumeration through the reference model."""
    import numpy as np
    sys.path.insert(0, "/root/reference")
    m = build_reference_model("retrieval", max_length=L)
    synth.load_synth(m, 0, "stress")
    imgs, caps = synth.synth_images(N, 4, 1.0), synth.synth_token_ids(N, L, 4)
    scores = torch.empty(N * N)
    with torch.no_grad():
        for i in range(N):
            scores[i * N:(i + 1) * N] = m(imgs[i:i + 1].expand(N, -1, -1, -1).contiguous(), caps)[:, 1]
    labels = torch.eye(N, dtype=torch.long)
    labels[1, 4] = labels[4, 1] = 1             # duplicate cap_id case, run_retrieval.py:139
    return {"task": "rank", "N": N, "L": L, "weight_seed": 0, "data_seed": 4, "scores": scores.view(N, N),
            "labels": labels}


def write_state_dict_keys():
    """tests/golden/state_dict_keys.json: key -> shape of the REAL reference's task models, in state_dict order."""
    import json
    out = {}
    for task in ("vqa", "retrieval", "pretrain", "caption"):
        out[task] = {k: list(v.shape) for k, v in build_reference_model(task).state_dict().items()}
    for conv in ("resnet101", "resnet50", "linear", "vit"):
        m = build_reference_model("retrieval", max_length=80, conv=conv)
        out[f"retrieval_{conv}"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(GOLDEN_DIR, "state_dict_keys.json"), "w") as f:
        json.dump(out, f)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if sys.argv[1:] == ["state_dict_keys"]:
        return write_state_dict_keys()
    torch.set_num_threads(os.cpu_count())
    cases = {
        "retrieval_stress": lambda: case_retrieval("stress", 1.0),
        "retrieval_config1": lambda: case_retrieval("init", 0.02),       # BASELINE.json configs[0]
        "vqa_stress": case_vqa,
        "pretrain_stress": case_pretrain,
        "rank6": case_rank,
        "retrieval_resnet101": lambda: case_retrieval("stress", 1.0, conv="resnet101"),   # BASELINE.json configs[4] backbone
        "retrieval_resnet50": lambda: case_retrieval("stress", 1.0, conv="resnet50"),
        "caption_stress": case_caption,
        "retrieval_linear": lambda: case_retrieval("stress", 1.0, conv="linear"),
        "retrieval_vit": lambda: case_retrieval("stress", 1.0, conv="vit"),
    }
    only = sys.argv[1:]                         # `python -m oracle.make_golden NAME...` regenerates just those cases
    for name, fn in cases.items():
        if only and name not in only:
            continue
        c = fn()
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(c, path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
