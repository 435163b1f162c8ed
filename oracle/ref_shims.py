"""Import harness for the UNMODIFIED reference (MVLT) — TEST INFRASTRUCTURE ONLY.

This file does not restate any arithmetic.  It only makes `/root/reference` importable in this
container so that

  * `oracle/make_golden.py` can run the real reference on seeded inputs and commit its outputs as
    fixtures under `tests/golden/`, and
  * `tests/test_oracle_vs_reference.py` can pin `oracle/mvlt_oracle.py` (the travelling CPU
    restatement) against the real thing whenever the reference tree is present.

`/root/reference` does not exist on the GPU box: nothing run there may import this module.

Shims (none touches the math):
  timm.models.layers   -> DropPath (identity in eval), to_2tuple, trunc_normal_   (visual_feature_extractor.py:122)
  yacs.config.CfgNode  -> attr-dict with merge/clone/freeze                       (swin_transformer_config.py:14)
  transformers.BeamSearchScorer -> placeholder (only report generation uses it)   (model.py:7)
  torch.load           -> {'model': {}} for the absent Swin .pth (strict=False)   (model.py:222-225)
  torchvision model_urls / load_state_dict_from_url -> offline random-init state_dict (ResNet variants)  (vfe.py:11)
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MVLT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "model.py"))


def _install_timm_stub():
    if "timm.models.layers" in sys.modules:
        return
    import torch.nn as nn

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            import torch
            keep = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            mask = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
            mask.floor_()
            return x.div(keep) * mask

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def _install_yacs_stub():
    if "yacs.config" in sys.modules:
        return
    import yaml

    class CfgNode(dict):
        def __init__(self, init=None, **kw):
            super().__init__()
            for k, v in (init or {}).items():
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            import copy
            return copy.deepcopy(self)

        def defrost(self):
            pass

        def freeze(self):
            pass

        def _merge(self, other):
            for k, v in other.items():
                if isinstance(v, dict) and isinstance(self.get(k), dict):
                    self[k]._merge(v)
                else:
                    self[k] = CfgNode(v) if isinstance(v, dict) else v

        def merge_from_file(self, path):
            with open(path) as f:
                self._merge(yaml.safe_load(f) or {})

        def merge_from_list(self, lst):
            for k, v in zip(lst[0::2], lst[1::2]):
                node = self
                parts = k.split(".")
                for p in parts[:-1]:
                    node = node[p]
                node[parts[-1]] = v

    yacs = types.ModuleType("yacs")
    config = types.ModuleType("yacs.config")
    config.CfgNode = CfgNode
    yacs.config = config
    sys.modules.update({"yacs": yacs, "yacs.config": config})


@contextlib.contextmanager
def _reference_cwd():
    """model.py:205/222 read a yaml and a .pth relative to CWD, and argparse reads sys.argv."""
    old_cwd, old_argv = os.getcwd(), sys.argv
    os.chdir(REFERENCE_ROOT)
    sys.argv = sys.argv[:1]
    try:
        yield
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv


_ref_modules = None


def import_reference():
    """Returns (modules.model, modules.config, modules.visual_feature_extractor) of the real reference."""
    global _ref_modules
    if _ref_modules is not None:
        return _ref_modules
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch
    # order matters: tokenizer import swaps the lazy module object, and
    # transformers probes for a real `timm` at import, so the stubs go in only afterwards
    from transformers import BertTokenizerFast, PreTrainedModel  # noqa: F401
    from transformers.models.bert.modeling_bert import BertEncoder  # noqa: F401
    import transformers
    _install_timm_stub()
    _install_yacs_stub()

    if not hasattr(sys.modules["transformers"], "BeamSearchScorer"):
        class BeamSearchScorer:  # placeholder; report generation is out of scope
            pass
        sys.modules["transformers"].BeamSearchScorer = BeamSearchScorer
        transformers.BeamSearchScorer = BeamSearchScorer

    if not getattr(torch.load, "_mvlt_shim", False):
        _orig_load = torch.load

        def _load(f, *a, **kw):
            if isinstance(f, str) and f.endswith("swin_small_patch4_window7_224.pth") and not os.path.exists(f):
                return {"model": {}}
            return _orig_load(f, *a, **kw)

        _load._mvlt_shim = True
        torch.load = _load

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with _reference_cwd():
        import modules.model as ref_model
        import modules.config as ref_config
        import modules.visual_feature_extractor as ref_vfe
    _ref_modules = (ref_model, ref_config, ref_vfe)
    return _ref_modules


def _install_resnet_shim(ref_vfe):
    """vfe.py:11 reads `torchvision.models.resnet.model_urls` (removed from torchvision >= 0.13) and downloads the ImageNet
    weights; model.py:196-199 hard-codes PRETRAIN=True.  There is no network: hand back a random-init state_dict (the
    caller overwrites every tensor with synthetic values anyway)."""
    import torchvision
    if not hasattr(torchvision.models.resnet, "model_urls"):
        torchvision.models.resnet.model_urls = {"resnet101": "offline://resnet101", "resnet50": "offline://resnet50"}

    def _offline_state_dict(url, *a, **kw):
        name = url.rsplit("/", 1)[-1]
        return getattr(torchvision.models, name)(weights=None).state_dict()

    if not hasattr(torchvision.models.vision_transformer, "model_urls"):       # vfe.py:84-87 (ViT variant)
        torchvision.models.vision_transformer.model_urls = {"vit_b_16": "offline://vit_b_16"}
    ref_vfe.load_state_dict_from_url = _offline_state_dict


def make_reference_config(task: str, max_length: int = 80, result_num: int = 224, itm: bool = True,
                          conv: str = "swintransformer"):
    """Config objects as the run_*.py scripts would build them, without network."""
    _, ref_config, _ = import_reference()
    cls = {"vqa": ref_config.MVLBertConfigforVQA, "retrieval": ref_config.MVLBertRetrieval,
           "pretrain": ref_config.MVLBertPretrainConfig, "caption": ref_config.MVLBertConfigForImageCaption}[task]
    cfg = cls()
    cfg.conv = conv
    cfg.vocab_size = 30522
    cfg.cls_token_id, cfg.sep_token_id, cfg.eos_token_id, cfg.mask_token_id = 101, 102, 104, 103
    cfg.max_length = max_length
    cfg.result_num = result_num
    if task == "pretrain":
        cfg.ITM_task = itm
    return cfg


def build_reference_model(task: str, seed: int = 0, **cfg_kw):
    """The real reference task model (eval mode), random init under torch.manual_seed(seed)."""
    import torch
    ref_model, _, ref_vfe = import_reference()
    cfg = make_reference_config(task, **cfg_kw)
    if cfg.conv.startswith("resnet") or cfg.conv.lower() in ("vit", "visiontransformer"):
        _install_resnet_shim(ref_vfe)
    cls = {"vqa": ref_model.MVLBertForVQA, "retrieval": ref_model.MVLBertForRetrieval,
           "pretrain": ref_model.MVLBertForPretraining, "caption": ref_model.MVLBertForImageCaption}[task]
    with _reference_cwd():
        torch.manual_seed(seed)
        model = cls(cfg)
    return model.eval()
