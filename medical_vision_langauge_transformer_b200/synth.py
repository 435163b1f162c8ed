"""Deterministic synthetic weights and inputs.

There is no network for datasets or checkpoints, so every parity test, golden fixture and bench run
draws its `state_dict` and inputs from here.  Generation is keyed on the tensor NAME (crc32) so the
same values come out for the reference modules (in the build container) and for this package's
modules (on the GPU box), independent of construction order.

Two weight flavours:
  * "stress"  – non-trivial LayerNorm gains/biases, non-zero Linear biases, O(1) attention logits and
                relative-position biases, so that a dropped bias / gamma / mask / index bug is visible
                even at bf16 tolerance.  Used for golden fixtures and parity tests.
  * "init"    – the distributions the reference constructs with (Swin trunc-normal 0.02 + zero bias +
                unit LN, vfe.py:659-666; BERT side PyTorch defaults because init_weights() is never
                called, model.py:365).  Used by bench.py ("random-init weights of that architecture").
"""
from __future__ import annotations

import math
import re
import zlib
from typing import Dict, Mapping

import torch

_BUFFERS = ("relative_position_index", "attn_mask", "position_ids", "num_batches_tracked")
_RESNET = re.compile(r"^conv\.conv\.0\.(conv1|bn1|layer[1-4]|fc)\.")      # torchvision ResNet keys under Conv_layer
_BATCHNORM = re.compile(r"\.(bn\d?|downsample\.1)\.(weight|bias|running_mean|running_var)$")
_VIT_LINEAR = re.compile(r"^conv\.conv\.0\.(class_token$|conv_proj\.|encoder\.|heads\.|linear_patch\.|bn\.)")   # vfe.py:47-107


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(name: str, shape, seed: int = 0, flavour: str = "stress") -> torch.Tensor:
    g = _gen(name, seed)
    shape = tuple(shape)
    randn = lambda s=1.0: torch.randn(shape, generator=g) * s
    leaf = name.rsplit(".", 1)[-1]
    if _RESNET.match(name) and ".fc." not in name:
        return _synth_resnet(name, shape, g, flavour)
    if _VIT_LINEAR.match(name):
        return _synth_vit_linear(name, shape, g, flavour)
    is_swin = name.startswith("conv.conv.0.")
    is_ln = any(t in name for t in ("norm", "LayerNorm")) and "downsample.reduction" not in name
    if is_ln:
        if flavour == "init":
            return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        return 1.0 + randn(0.1) if leaf == "weight" else randn(0.05)
    if "relative_position_bias_table" in name:
        return randn(0.02).clamp_(-0.04, 0.04) if flavour == "init" else randn(0.3)
    if "embeddings.weight" in name:
        return randn(1.0)                                   # nn.Embedding default N(0,1)
    if leaf == "bias":
        if flavour == "init" and is_swin and "patch_embed.proj" not in name:
            return torch.zeros(shape)
        return randn(0.02)
    if leaf == "weight" and len(shape) >= 2:
        fan_in = int(math.prod(shape[1:]))
        if is_swin and "patch_embed.proj" not in name:
            if flavour == "init":
                return randn(0.02).clamp_(-0.04, 0.04)
            # qkv scaled so window-attention logits are O(1) instead of ~0 (sharper softmax)
            return randn(0.8 / math.sqrt(fan_in)) if ".attn.qkv." in name else randn(0.5 / math.sqrt(fan_in))
        bound = 1.0 / math.sqrt(fan_in)                     # nn.Linear / Conv2d default
        if flavour == "stress" and ".attention.self." in name and "value" not in name:
            bound *= 2.0
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    return randn(0.02)


def _synth_resnet(name: str, shape, g: torch.Generator, flavour: str) -> torch.Tensor:
    """ResNet trunk (vfe.py:7-44).  "init": torchvision's construction (kaiming-normal fan_out convolutions, unit BatchNorm,
    fresh running statistics).  "stress": fan_in He convolutions and non-trivial BatchNorm parameters / running statistics;
    the last BatchNorm of each bottleneck has a gain around 0.25 so 33 residual blocks stay O(10) instead of doubling the
    variance per block (which only tests the exponent range)."""
    leaf = name.rsplit(".", 1)[-1]
    randn = lambda s=1.0: torch.randn(shape, generator=g) * s
    if _BATCHNORM.search(name):
        if flavour == "init":
            return torch.ones(shape) if leaf in ("weight", "running_var") else torch.zeros(shape)
        if leaf == "weight":
            return (0.25 + randn(0.025)) if ".bn3." in name else (1.0 + randn(0.1))
        if leaf == "running_var":
            return 0.8 + 0.4 * torch.rand(shape, generator=g)
        return randn(0.05)
    assert leaf == "weight" and len(shape) == 4, name
    n, c, r, s = shape
    return randn(math.sqrt(2.0 / (n * r * s))) if flavour == "init" else randn(math.sqrt(2.0 / (c * r * s)))


def _synth_vit_linear(name: str, shape, g: torch.Generator, flavour: str) -> torch.Tensor:
    """torchvision ViT-B/16 trunk (vfe.py:66-107) and the linear-patch stem (vfe.py:47-60).  "init": unit LayerNorm / BatchNorm,
    zero biases, small normal weights; "stress": non-trivial norms and biases, O(1) attention logits, damped residual
    branches (out_proj / mlp.3) so twelve pre-LN blocks stay O(1)."""
    leaf = name.rsplit(".", 1)[-1]
    randn = lambda s=1.0: torch.randn(shape, generator=g) * s
    init = flavour == "init"
    if _BATCHNORM.search(name):
        if init:
            return torch.ones(shape) if leaf in ("weight", "running_var") else torch.zeros(shape)
        if leaf == "weight":
            return 1.0 + randn(0.1)
        return 0.8 + 0.4 * torch.rand(shape, generator=g) if leaf == "running_var" else randn(0.05)
    if ".ln_1." in name or ".ln_2." in name or ".encoder.ln." in name:
        if init:
            return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        return 1.0 + randn(0.1) if leaf == "weight" else randn(0.05)
    if leaf == "class_token":
        return torch.zeros(shape) if init else randn(0.5)
    if leaf == "pos_embedding":
        return randn(0.02) if init else randn(0.2)
    if leaf in ("bias", "in_proj_bias"):
        return torch.zeros(shape) if init else randn(0.02)
    fan_in = int(math.prod(shape[1:]))
    if leaf == "in_proj_weight":
        return randn(0.03) if init else randn(1.5 / math.sqrt(fan_in))
    damp = 0.5 if (".out_proj." in name or ".mlp.3." in name) and not init else 1.0
    return randn(damp / math.sqrt(fan_in))


def synth_state_dict(shapes: Mapping[str, torch.Size], seed: int = 0, flavour: str = "stress") -> Dict[str, torch.Tensor]:
    """New fp32 tensors for every parameter in `shapes` (buffers listed in _BUFFERS are skipped)."""
    out = {}
    for name, shape in shapes.items():
        if any(name.endswith(b) for b in _BUFFERS):
            continue
        out[name] = synth_tensor(name, shape, seed, flavour)
    return out


def load_synth(model: torch.nn.Module, seed: int = 0, flavour: str = "stress") -> Dict[str, torch.Tensor]:
    """Overwrite every parameter of `model` in place with the synthetic values; returns the full state_dict."""
    sd = model.state_dict()
    new = synth_state_dict({k: v.shape for k, v in sd.items()}, seed, flavour)
    with torch.no_grad():
        for k, v in new.items():
            sd[k].copy_(v.to(sd[k].dtype))
    return model.state_dict()


def synth_images(batch: int, seed: int = 1, scale: float = 0.02) -> torch.Tensor:
    """[B,3,224,224] fp32.  preprocess_rgc.py:38-39 divides by the VARIANCE, hence |x| <~ 0.05 -> randn*0.02;
    scale=1.0 is the stress case."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    return torch.randn(batch, 3, 224, 224, generator=g) * scale


def synth_token_ids(batch: int, max_length: int = 80, seed: int = 1, min_len: int = 10) -> torch.Tensor:
    """int64 [B,L]: ids uniform in [1000,30000), last real token [END]=104, zero ([PAD]) suffix
    (run_retrieval.py:141-145, run_pretrain.py:119-122)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 3)
    ids = torch.randint(1000, 30000, (batch, max_length), generator=g)
    lo = min(min_len, max_length)
    lens = torch.randint(lo, max_length + 1, (batch,), generator=g)
    pos = torch.arange(max_length)[None]
    ids = torch.where(pos < lens[:, None], ids, torch.zeros_like(ids))
    ids[torch.arange(batch), lens - 1] = 104
    return ids


def synth_mlm_labels(ids: torch.Tensor, seed: int = 1, max_masked: int = 10):
    """Masked caption + labels as run_pretrain.py:130-158 shapes them: <=10 positions per sample carry the
    original id as label and [MASK]=103 as input; every other label is -100."""
    g = torch.Generator(device="cpu")
    g.manual_seed(104729 * seed + 5)
    masked, labels = ids.clone(), torch.full_like(ids, -100)
    for b in range(ids.shape[0]):
        n = int((ids[b] > 0).sum())
        k = max(1, min(max_masked, n // 7))
        pos = torch.randperm(n, generator=g)[:k]
        labels[b, pos] = ids[b, pos]
        masked[b, pos] = 103
    return masked, labels
