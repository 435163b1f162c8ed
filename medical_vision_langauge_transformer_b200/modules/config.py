"""Config classes with the names and fields of the reference's modules/config.py:4-72.

They are plain `BertConfig` subclasses (so `from_pretrained` / `save_pretrained` directories written by the reference
load unchanged); defaults are bert-base (768/12/12/3072, vocab 30522, LayerNorm eps 1e-12, erf GELU)."""
from __future__ import annotations

from transformers.models.bert.configuration_bert import BertConfig

_SPECIAL = ("[END]", "[CLS]", "[SEP]", "[MASK]")


class MVLBertConfig(BertConfig):
    """Base config (reference config.py:4-27)."""

    _task_defaults = {}

    def __init__(self, **kwargs):
        mlm, itm = kwargs.pop("MLM_task", True), kwargs.pop("ITM_task", True)
        conv, max_length = kwargs.pop("conv", "resnet101"), kwargs.pop("max_length", 40)
        super().__init__(**kwargs)
        self.type_vocab_size = 3                       # 0 text, 1 image (model.py:152-153)
        self.MLM_task, self.ITM_task = mlm, itm
        self.conv = conv
        self.result_num = 224
        self.lr = 4e-5
        self.max_length = max_length
        self.mask_token_id = None
        self.attention_probs_dropout_prob = 0.0
        self.hidden_dropout_prob = 0.0
        for k, v in type(self)._task_defaults.items():
            setattr(self, k, v)

    def update_special_tokens(self, tokenizer):
        ids = tokenizer.convert_tokens_to_ids(list(_SPECIAL))
        self.eos_token_id, self.cls_token_id, self.sep_token_id, self.mask_token_id = ids
        self.vocab_size = len(tokenizer)
        print("eos_token_id:", self.eos_token_id, "cls_token_id:", self.cls_token_id, "sep_token_id:", self.sep_token_id,
              "mask_token_id:", self.mask_token_id)


class MVLBertConfigforVQA(MVLBertConfig):
    """reference config.py:29-38"""
    _task_defaults = dict(MLM_task=True, ITM_task=True, result_num=224, lr=4e-5, attention_probs_dropout_prob=0.1,
                          hidden_dropout_prob=0.1)

    def __init__(self, **kwargs):   # explicit: HF 5.x turns config classes into dataclasses and would synthesise one
        super().__init__(**kwargs)


class MVLBertPretrainConfig(MVLBertConfig):
    """reference config.py:41-50"""
    _task_defaults = dict(MLM_task=True, ITM_task=False, max_length=150, lr=4e-5, attention_probs_dropout_prob=0.1,
                          hidden_dropout_prob=0.1)

    def __init__(self, **kwargs):   # explicit: HF 5.x turns config classes into dataclasses and would synthesise one
        super().__init__(**kwargs)


class MVLBertRetrieval(MVLBertConfig):
    """reference config.py:53-60"""
    _task_defaults = dict(ITM_task=True, lr=1e-6, max_length=80, attention_probs_dropout_prob=0.1)

    def __init__(self, **kwargs):   # explicit: HF 5.x turns config classes into dataclasses and would synthesise one
        super().__init__(**kwargs)


class MVLBertConfigForImageCaption(MVLBertConfig):
    """reference config.py:64-72 (report generation; the teacher-forced pass runs natively, the decode loop does not)."""
    _task_defaults = dict(lr=1e-5, max_length=80, is_decoder=True, attention_probs_dropout_prob=0.1,
                          hidden_dropout_prob=0.1)

    def __init__(self, **kwargs):   # explicit: HF 5.x turns config classes into dataclasses and would synthesise one
        super().__init__(**kwargs)


def offline_config(task: str, conv: str = "swintransformer", max_length: int = 80, **overrides) -> MVLBertConfig:
    """What `Config.from_pretrained('bert-base-uncased')` + `update_special_tokens(tokenizer)` yields, without network:
    bert-base defaults, vocab 30522, [CLS]=101 [SEP]=102 [MASK]=103 [END]=104 (dataset/bert-base-uncased/vocab.txt)."""
    cls = {"vqa": MVLBertConfigforVQA, "retrieval": MVLBertRetrieval, "pretrain": MVLBertPretrainConfig,
           "caption": MVLBertConfigForImageCaption}[task]
    cfg = cls()
    cfg.conv = conv
    cfg.vocab_size = 30522
    cfg.cls_token_id, cfg.sep_token_id, cfg.mask_token_id, cfg.eos_token_id = 101, 102, 103, 104
    cfg.max_length = max_length
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg
