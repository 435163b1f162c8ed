"""MVLT joint image-text encoder and task heads with the class names, constructor arguments, state_dict keys and
forward signatures of the reference's modules/model.py:16-476, executed by libmvlt_b200.so.

Parameter-holder classes (`_BertLayer` ...) reproduce the HuggingFace BertEncoder/BertPooler/heads key layout
(`MVLBert.encoder.layer.{l}.attention.self.query.weight`, ...) so reference checkpoints — `save_pretrained`
directories and whole-model pickles — load unchanged.  Weights keep PyTorch-default initialisation exactly as in
the reference, where `init_weights()` is never called (model.py:365).

Out of scope here: the KV-cache decode branch of `get_embedding` (model.py:82-108) and
the greedy / beam-search decode of `MVLBertForImageCaption` (its teacher-forced pass is here).
"""
from __future__ import annotations

import os

import random

import torch
import torch.nn as nn
from transformers import PreTrainedModel

from .. import ops
from .config import MVLBertConfig
from .visual_feature_extractor import (SwinTransformer, VisionTransformerBaseWithoutPooling, act_dtype, default_precision,
                                       linear_patch_16x16, resnet50_without_poolfc, resnet101_without_fc)

# Swin-S, the only backbone config the reference ships enabled (modules/swin_small_patch4_window7_224.yaml:1-8 on top
# of the defaults in swin_transformer_config.py); `Conv_layer(config, swin_kwargs=...)` overrides it.
SWIN_SMALL = dict(img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 18, 2],
                  num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0,
                  drop_path_rate=0.3, ape=False, patch_norm=True, use_checkpoint=False)


# ------------------------------------------------------------------------------------------------ holders (HF layout)
class _SelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.query, self.key, self.value = (nn.Linear(c.hidden_size, c.hidden_size) for _ in range(3))


class _DenseLN(nn.Module):
    """BertSelfOutput / BertOutput / BertPredictionHeadTransform key layout: dense + LayerNorm."""

    def __init__(self, c, d_in):
        super().__init__()
        self.dense = nn.Linear(d_in, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _Attention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = _SelfAttention(c)
        self.output = _DenseLN(c, c.hidden_size)


class _Intermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)


class _BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _Attention(c)
        self.intermediate = _Intermediate(c)
        self.output = _DenseLN(c, c.intermediate_size)


class _BertEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(c) for _ in range(c.num_hidden_layers)])


class _Pooler(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)


class _LMPredictionHead(nn.Module):
    """HF modeling_bert.py:488-503 keys: transform.{dense,LayerNorm}, decoder.{weight,bias}, bias (unused in forward)."""

    def __init__(self, c):
        super().__init__()
        self.transform = _DenseLN(c, c.hidden_size)
        self.decoder = nn.Linear(c.hidden_size, c.vocab_size, bias=True)
        self.bias = nn.Parameter(torch.zeros(c.vocab_size))


class _OnlyMLMHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.predictions = _LMPredictionHead(c)


def _require_gelu(config):
    if getattr(config, "hidden_act", "gelu") != "gelu":
        raise NotImplementedError(f"hidden_act={config.hidden_act!r}: the fused epilogues implement exact erf GELU only")


class _PackedMixin:
    """Lazy cache of kernel-layout weights keyed on (precision, device, sum of parameter versions)."""

    def _pack_params(self):
        return self.parameters()

    def _fingerprint(self):
        ps = list(self._pack_params())
        return (self.precision, ps[0].device, id(ps[0]), sum(p._version for p in ps))

    def packed(self):
        key = self._fingerprint()
        if getattr(self, "_pk", None) is None or self._pk_key != key:
            self._pk, self._pk_key = self._pack(), key
        return self._pk


# ------------------------------------------------------------------------------------------------ MVLBert
class MVLBert(_PackedMixin, nn.Module):
    """model.py:16-183."""

    def __init__(self, config, add_pooling_layer=False, precision=None):
        super().__init__()
        _require_gelu(config)
        self.config_class = MVLBertConfig
        self.config = config
        self.precision = precision or default_precision()
        self.word_embeddings = nn.Embedding(config.vocab_size + 1, config.hidden_size)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.embedding_LayerNorm = nn.LayerNorm(config.hidden_size, eps=1e-12)   # declared, never applied (model.py:25,158)
        self.embedding_dropout = nn.Dropout(config.hidden_dropout_prob)
        self.encoder = _BertEncoder(config)
        self.is_decoder = config.is_decoder
        self.pooler = _Pooler(config) if add_pooling_layer else None
        self.register_buffer("position_ids", torch.arange(512).expand((1, -1)))
        self.taps = None   # tests set this to a dict to record activations (clones)

    def _pack(self):
        wd = act_dtype(self.precision)
        f32 = lambda t: t.detach().float().contiguous()
        wcast = lambda t: t.detach().to(wd).contiguous()
        layers = []
        for l in self.encoder.layer:
            sa = l.attention.self
            layers.append(dict(
                qkv_w=wcast(torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0)),
                qkv_b=f32(torch.cat([sa.query.bias, sa.key.bias, sa.value.bias], 0)),
                ao_w=wcast(l.attention.output.dense.weight), ao_b=f32(l.attention.output.dense.bias),
                ln1_w=f32(l.attention.output.LayerNorm.weight), ln1_b=f32(l.attention.output.LayerNorm.bias),
                fi_w=wcast(l.intermediate.dense.weight), fi_b=f32(l.intermediate.dense.bias),
                fo_w=wcast(l.output.dense.weight), fo_b=f32(l.output.dense.bias),
                ln2_w=f32(l.output.LayerNorm.weight), ln2_b=f32(l.output.LayerNorm.bias)))
        pk = dict(layers=layers, word=f32(self.word_embeddings.weight))
        if self.pooler is not None:
            pk["pool_w"], pk["pool_b"] = wcast(self.pooler.dense.weight), f32(self.pooler.dense.bias)
        return pk

    def _typepos(self, pk, n_obj, S):
        """token_type_emb[s <= obj_end] + position_emb[s]  (model.py:152-157) — input independent, cached per (n_obj,S)."""
        key = ("typepos", n_obj, S)
        if key not in pk:
            pos = torch.arange(S, device=self.position_ids.device)
            typ = (pos <= n_obj + 1).long()
            pk[key] = (self.token_type_embeddings.weight.detach().float()[typ]
                       + self.position_embeddings.weight.detach().float()[self.position_ids[0, :S]]).contiguous()
        return pk[key]

    def encode(self, text_idx, text_mask, image_feature, image_mask=None, seq2seq_mask=False, img_index=None):
        """Embedding + 12 post-LN layers.  -> (hidden fp32 [B*S, D], its bf16 shadow or None, B, S).

        The hidden state / residual stays fp32 in both precisions; in bf16 mode every LayerNorm also emits the rows
        rounded to bf16 as the A operand of the next tcgen05 GEMM (the residual path never sees bf16 rounding)."""
        cfg = self.config
        if self.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0):
            raise NotImplementedError("train mode (Dropout, HF modeling_bert.py:296,:353) is outside the accelerated forward path; "
                                      "call model.eval() — this package implements the eval-mode forward only")
        pk = self.packed()
        bf = self.precision == "bf16"
        image_feature = image_feature.float().contiguous()
        n_obj = image_feature.shape[1]
        B, L = text_idx.shape
        S = n_obj + 2 + L
        if S > cfg.max_position_embeddings:
            raise ValueError(f"joint sequence {S} exceeds max_position_embeddings {cfg.max_position_embeddings}")
        h, hb, kmask = ops.joint_embed(image_feature, text_idx.contiguous(), text_mask, image_mask, pk["word"],
                                       self._typepos(pk, n_obj, S), cfg.cls_token_id, cfg.sep_token_id, img_index,
                                       bf16_copy=bf)
        heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
        taps = self.taps
        # MVLT_LINEAR_LN=0 keeps the unfused GEMM (reduce-add epilogue) + layernorm_rows chain (bf16 mode; fp32 mode always does)
        fused_ln = bf and cfg.hidden_size in ops.LINEAR_LN_WIDTHS and os.environ.get("MVLT_LINEAR_LN", "1") != "0" and ops.use_linear_ln(B * S)
        if taps is not None:
            taps["image_feature"] = image_feature.clone()
            taps["embedding"] = h.clone().view(B, S, -1)
        for li, w in enumerate(pk["layers"]):
            qkv = ops.linear(hb if bf else h, w["qkv_w"], w["qkv_b"])
            ctx = ops.joint_attention(qkv, kmask, B, S, heads, bool(seq2seq_mask), n_obj + 1)
            # residual GEMMs accumulate IN PLACE into the fp32 hidden state (the tcgen05 epilogue reduce-adds its tile at the
            # L2, the residual never enters the SM); each LayerNorm then rewrites the rows it has just read
            if fused_ln:
                # dense + residual + LayerNorm as ONE tcgen05 kernel on clusters of four CTAs (csrc/gemm_ln.cu): the fp32 rows are
                # read once and written once per half-layer instead of GEMM read-modify-write + LayerNorm read + write
                h1, h1b = ops.linear_residual_layernorm(ctx, w["ao_w"], w["ao_b"], h, w["ln1_w"], w["ln1_b"], eps, out=h)
                f = ops.linear(h1b, w["fi_w"], w["fi_b"], act=ops.ACT_GELU)
                h, hb = ops.linear_residual_layernorm(f, w["fo_w"], w["fo_b"], h1, w["ln2_w"], w["ln2_b"], eps, out=h1)
            else:
                ops.linear(ctx, w["ao_w"], w["ao_b"], residual=h, out=h)
                h1 = ops.layernorm(h, w["ln1_w"], w["ln1_b"], eps, torch.float32, out=h, bf16_copy=bf)
                h1, h1b = h1 if bf else (h1, None)
                f = ops.linear(h1b if bf else h1, w["fi_w"], w["fi_b"], act=ops.ACT_GELU)
                ops.linear(f, w["fo_w"], w["fo_b"], residual=h1, out=h1)
                h = ops.layernorm(h1, w["ln2_w"], w["ln2_b"], eps, torch.float32, out=h1, bf16_copy=bf)
                h, hb = h if bf else (h, None)
            if taps is not None and li in (0, len(pk["layers"]) - 1):
                taps[f"bert{li}"] = h.clone().view(B, S, -1)
        return h, hb, B, S

    def pool(self, hidden, shadow, B, S):
        """BertPooler (HF modeling_bert.py:456-468): tanh(W h[:,0] + b) over the B [CLS] rows (row stride S*D).
        bf16 mode: tcgen05 GEMM on the bf16 shadow rows, pooled returned in bf16 (it only feeds further GEMMs);
        fp32 mode: CUDA-core GEMM on the fp32 rows."""
        pk = self.packed()
        if shadow is not None:
            pooled = ops.linear(shadow.view(B, S, -1)[:, 0], pk["pool_w"], pk["pool_b"], act=ops.ACT_TANH, out_dtype=torch.bfloat16)
        else:
            pooled = ops.linear(hidden.view(B, S, -1)[:, 0], pk["pool_w"], pk["pool_b"], act=ops.ACT_TANH)
        if self.taps is not None:
            self.taps["pooled"] = pooled.float()
        return pooled

    def forward(self, text_idx, text_mask, image_feature, image_mask, past_key_values=None, use_cache=False,
                seq2seq_mask=False, output_text_image_seperate=False):
        if past_key_values is not None or use_cache:
            raise NotImplementedError("KV-cache decoding (model.py:82-108) is outside the accelerated forward path")
        if text_idx is None:
            raise NotImplementedError("text_idx=None (generation warm-up, model.py:145-147) is outside the accelerated path")
        hidden, shadow, B, S = self.encode(text_idx, text_mask, image_feature, image_mask, seq2seq_mask)
        last = hidden.view(B, S, -1)
        pooled = self.pool(hidden, shadow, B, S) if self.pooler is not None else None
        if pooled is not None and pooled.dtype != torch.float32:
            pooled = pooled.float()
        if output_text_image_seperate:
            obj_end = image_feature.shape[1] + 1
            text_end = obj_end + text_idx.shape[1] + 1
            return last[:, obj_end + 1:text_end], last[:, 1:obj_end], pooled, last[:, obj_end]
        return (last,), pooled


# ------------------------------------------------------------------------------------------------ Conv_layer
class Conv_layer(nn.Module):
    """model.py:186-266: Sequential(backbone, GELU) -> [B, 49, 768], returned in fp32 (as in the reference) in both
    precision modes.  Swin branch: the final LayerNorm and the GELU run as one kernel.  ResNet-101 / ResNet-50 branch
    (model.py:195-201): the GELU is the epilogue of the last bottleneck's GEMM, then `[B, 2048, 7, 7] -> [B, 49, 2048]`
    (a no-op for NHWC activations) and `resnet_fc` (model.py:236, :263-264).  `linear` (model.py:200-201) and `vit` /
    `visiontransformer` (:227-228): 196 image tokens of width 768, GELU in the last kernel's epilogue."""

    def __init__(self, config, swin_kwargs=None, precision=None):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size if config is not None else 768
        kind = str(config.conv).lower()
        if str(config.conv) == "resnet101":
            backbone = resnet101_without_fc(precision=precision)     # reference: ImageNet weights by URL (no network here)
        elif str(config.conv) == "resnet50":
            backbone = resnet50_without_poolfc(precision=precision)
        elif kind == "swintransformer":
            kw = dict(SWIN_SMALL)
            kw.update(swin_kwargs or {})
            backbone = SwinTransformer(precision=precision, **kw)
        elif str(config.conv) == "linear":
            backbone = linear_patch_16x16(precision=precision)
        elif kind in ("vit", "visiontransformer"):
            backbone = VisionTransformerBaseWithoutPooling(precision=precision)      # reference: ImageNet weights by URL
        else:
            raise NotImplementedError("no such config.conv")
        self.conv = nn.Sequential(backbone, nn.GELU())
        self.resnet_fc = nn.Linear(2048, config.hidden_size)     # unused on the Swin branch; kept for state_dict parity
        self._fc_pk = None

    def _fc_packed(self, precision):
        w, b = self.resnet_fc.weight, self.resnet_fc.bias
        key = (precision, w.device, w._version + b._version, id(w))
        if self._fc_pk is None or self._fc_pk[0] != key:
            self._fc_pk = (key, w.detach().to(act_dtype(precision)).contiguous(), b.detach().float().contiguous())
        return self._fc_pk[1], self._fc_pk[2]

    def forward(self, v):
        if torch.is_tensor(v) and v.dim() == 5:
            # IU-Xray: [batch, 2, channel, h, w] -> the two views' objects concatenated, [batch, 2*49, hidden] (model.py:240-253,
            # then resnet_fc per object, :263-264).  Both views go through the backbone as ONE batch of 2B images.
            B = v.shape[0]
            obj = self.forward(v.transpose(0, 1).reshape(2 * B, *v.shape[2:]))            # [2B, 49, D], view-major
            return torch.cat((obj[:B], obj[B:]), dim=1)
        backbone = self.conv[0]
        if isinstance(backbone, SwinTransformer):
            return backbone.forward_features(v, final_gelu=True, out_dtype=torch.float32)
        if isinstance(backbone, VisionTransformerBaseWithoutPooling):
            return backbone.forward_features(v, final_gelu=True)     # [B, 196, 768] (3-D: no reshape, width 768: no resnet_fc)
        if isinstance(backbone, linear_patch_16x16):
            a, H, W = backbone.forward_features(v, final_gelu=True)  # [B*196, 768] = the reshape/transpose of model.py:258-261
            return a.view(v.shape[0], H * W, -1)
        a, H, W = backbone.forward_features(v, final_gelu=True)      # [B*49, 2048], GELU applied
        fw, fb = self._fc_packed(backbone.precision)
        feat = ops.linear(a, fw, fb, out_dtype=torch.float32)        # model.py:263-264 (channel == 2048)
        return feat.view(v.shape[0], H * W, -1)


# ------------------------------------------------------------------------------------------------ task models
class MVLBertPretrainedModel(PreTrainedModel):
    """model.py:269-294."""
    base_model_prefix = "MVLBert"
    config_class = MVLBertConfig
    _keys_to_ignore_on_load_missing = [r"position_ids"]

    _constructing = False

    def __init__(self, config, *args, **kwargs):
        # Fresh construction keeps PyTorch-default initialisation exactly as in the reference, where `init_weights()` is never
        # called (model.py:365): `_init_weights` is inert until `post_init()` (end of every task `__init__`) has run.
        self._constructing = True
        super().__init__(config, *args, **kwargs)

    def _finish_init(self):
        self.post_init()
        self._constructing = False

    def _init_weights(self, module):
        """model.py:280-294, applied by `from_pretrained` to the tensors a checkpoint does NOT provide (HF builds the model on a
        meta device, so whatever is not loaded has no values): Linear / Embedding ~ N(0, initializer_range), LayerNorm = (1, 0).
        Beyond the reference: the input-independent buffers and the non-Linear parameters get their constructor values back.
        Every initialiser below skips tensors already loaded (`_is_hf_initialized`)."""
        if self._constructing:
            return
        from transformers import initialization as init
        from .visual_feature_extractor import SwinTransformerBlock, WindowAttention
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            init.normal_(module.weight, mean=0.0, std=std)
            if module.bias is not None:
                init.zeros_(module.bias)
        elif isinstance(module, nn.Embedding):
            init.normal_(module.weight, mean=0.0, std=std)
        elif isinstance(module, nn.LayerNorm):
            init.zeros_(module.bias)
            init.ones_(module.weight)
        elif isinstance(module, nn.Conv2d):
            init.kaiming_uniform_(module.weight, a=5 ** 0.5)
            if module.bias is not None:
                init.zeros_(module.bias)
        elif isinstance(module, nn.BatchNorm2d):
            init.ones_(module.weight); init.zeros_(module.bias)
            init.zeros_(module.running_mean); init.ones_(module.running_var); init.zeros_(module.num_batches_tracked)
        elif isinstance(module, WindowAttention):
            init.trunc_normal_(module.relative_position_bias_table, std=0.02)
            init.copy_(module.relative_position_index, module.build_relative_position_index())
        elif isinstance(module, SwinTransformerBlock):
            if module.attn_mask is not None:
                init.copy_(module.attn_mask, module.build_attn_mask())
        elif isinstance(module, MVLBert):
            init.copy_(module.position_ids, torch.arange(module.position_ids.shape[-1]).expand((1, -1)))
        elif isinstance(module, _LMPredictionHead):
            init.zeros_(module.bias)
        else:
            for prm in module.parameters(recurse=False):      # class tokens / position embeddings of the ViT variant
                init.normal_(prm, mean=0.0, std=std)

    @property
    def precision(self):
        return self.MVLBert.precision

    def set_precision(self, precision: str):
        """'bf16' (tcgen05 path) or 'fp32' (CUDA-core parity path)."""
        assert precision in ("bf16", "fp32")
        self.MVLBert.precision = precision
        self.conv.conv[0].precision = precision
        return self

    def _trunk(self, image, text_idx, image_mask, seq2seq=False):
        feat = self.conv(image)
        hidden, shadow, B, S = self.MVLBert.encode(text_idx, None, feat, image_mask, seq2seq)  # text_mask = ids>0 in-kernel
        return feat, hidden, shadow, B, S


class MVLBertForVQA(_PackedMixin, MVLBertPretrainedModel):
    """model.py:297-349: conv -> MVLBert -> Dropout(eval: identity) + Linear(768, result_num) -> softmax."""

    def __init__(self, config):
        super().__init__(config)
        self.config = config
        self.conv = Conv_layer(config)
        self.activation = nn.GELU()
        self.add_pooling_layer = True
        self.MVLBert = MVLBert(config, add_pooling_layer=True)
        self.final_mlp = nn.Sequential(nn.Dropout(config.hidden_dropout_prob, inplace=False),
                                       nn.Linear(config.hidden_size, config.result_num))
        self.softmax = nn.Softmax(dim=-1)
        self._finish_init()

    def _pack_params(self):
        return self.final_mlp.parameters()

    def _pack(self):
        lin = self.final_mlp[1]
        return dict(w=lin.weight.detach().to(act_dtype(self.precision)).contiguous(), b=lin.bias.detach().float().contiguous())

    def forward(self, image, question, label, image_mask=None):
        _, hidden, shadow, B, S = self._trunk(image, question, image_mask)
        pooled = self.MVLBert.pool(hidden, shadow, B, S)
        pk = self.packed()
        logits = ops.linear(pooled, pk["w"], pk["b"], out_dtype=torch.float32)     # B x result_num
        return ops.softmax_rows(logits), logits


class MVLBertForRetrieval(_PackedMixin, MVLBertPretrainedModel):
    """model.py:423-476: conv -> MVLBert -> BertPredictionHeadTransform + Linear(768,2) -> softmax(dim=1) | logits."""

    def __init__(self, config):
        super().__init__(config)
        self.config = config
        self.conv = Conv_layer(config)
        self.MVLBert = MVLBert(config, add_pooling_layer=True)
        self.final_mlp = nn.Sequential(_DenseLN(config, config.hidden_size), nn.Linear(config.hidden_size, 2))
        self.softmax = nn.Softmax(dim=1)
        self.sigmoid = nn.Sigmoid()
        self._finish_init()

    def _pack_params(self):
        return self.final_mlp.parameters()

    def _pack(self):
        t, lin = self.final_mlp[0], self.final_mlp[1]
        f32 = lambda x: x.detach().float().contiguous()
        return dict(tw=t.dense.weight.detach().to(act_dtype(self.precision)).contiguous(), tb=f32(t.dense.bias),
                    lw=f32(t.LayerNorm.weight), lb=f32(t.LayerNorm.bias), w=f32(lin.weight), b=f32(lin.bias))

    def head_logits(self, pooled):
        """BertPredictionHeadTransform (HF :471-485) + Linear(768,2) -> fp32 logits [B,2]."""
        pk = self.packed()
        t = ops.linear(pooled, pk["tw"], pk["tb"], act=ops.ACT_GELU, out_dtype=torch.float32)   # one row per pair
        t = ops.layernorm(t, pk["lw"], pk["lb"], self.config.layer_norm_eps, torch.float32)
        return ops.linear_small(t, pk["w"], pk["b"])

    def forward(self, image, caption, image_text_label=None, image_mask=None):
        _, hidden, shadow, B, S = self._trunk(image, caption, image_mask)
        logits = self.head_logits(self.MVLBert.pool(hidden, shadow, B, S))
        if image_text_label is None:
            return ops.softmax_rows(logits)
        return logits


class MVLBertForPretraining(_PackedMixin, MVLBertPretrainedModel):
    """model.py:352-420: MLM over the text positions (seq2seq or bidirectional mask, chosen by random.random() as in
    the reference) + optional ITM; returns the scalar loss."""

    def __init__(self, config):
        super().__init__(config)
        self.config = config
        self.config.output_text_and_image_seperately = True
        self.conv = Conv_layer(config)
        self.MVLBert = MVLBert(config, add_pooling_layer=True)
        self.MLM_head_seq2seq = _OnlyMLMHead(config)
        self.MLM_head_bidir = _OnlyMLMHead(config)
        self.ITM_mlp = nn.Linear(config.hidden_size, 2)
        self._finish_init()

    def _pack_params(self):
        return [*self.MLM_head_seq2seq.parameters(), *self.MLM_head_bidir.parameters(), *self.ITM_mlp.parameters()]

    def _pack(self):
        wd = act_dtype(self.precision)
        f32 = lambda x: x.detach().float().contiguous()
        out = {}
        for name, head in (("seq2seq", self.MLM_head_seq2seq), ("bidir", self.MLM_head_bidir)):
            p = head.predictions
            out[name] = dict(tw=p.transform.dense.weight.detach().to(wd).contiguous(), tb=f32(p.transform.dense.bias),
                             lw=f32(p.transform.LayerNorm.weight), lb=f32(p.transform.LayerNorm.bias),
                             dw=p.decoder.weight.detach().to(wd).contiguous(), db=f32(p.decoder.bias))
        out["itm_w"], out["itm_b"] = f32(self.ITM_mlp.weight), f32(self.ITM_mlp.bias)
        return out

    def forward(self, image, caption_masked, caption_label, image_text_label, image_mask=None):
        cfg = self.config
        seq2seq = random.random() < 0.5                                   # model.py:390-394: one draw per call
        feat, hidden, shadow, B, S = self._trunk(image, caption_masked, image_mask, seq2seq)
        adt = act_dtype(self.precision)
        n_obj, L, D = feat.shape[1], caption_masked.shape[1], hidden.shape[1]
        pk = self.packed()
        mlm_loss = torch.zeros((1, 1), device=hidden.device)     # model.py:405 (a CUDA tensor there as well once .cuda()'d)
        if cfg.MLM_task:
            w = pk["seq2seq" if seq2seq else "bidir"]
            src = shadow if shadow is not None else hidden                                # GEMM A operand dtype
            text = src.view(B, S, D)[:, n_obj + 2:n_obj + 2 + L].reshape(B * L, D)        # rows 51..51+L of each sample
            t = ops.linear(text, w["tw"], w["tb"], act=ops.ACT_GELU, out_dtype=torch.float32)
            t = ops.layernorm(t, w["lw"], w["lb"], cfg.layer_norm_eps, adt)
            V = w["dw"].shape[0]
            if self.precision == "bf16":
                # vocabulary GEMM with an online-logsumexp epilogue: the [B*L, vocab] fp32 logits (312 MB at batch 32) are never written
                acc = ops.mlm_ce_fused(t, w["dw"], w["db"], caption_label.reshape(-1), -100)
            else:
                ld = (V + 31) // 32 * 32
                logits = torch.empty((B * L, ld), device=t.device, dtype=torch.float32)[:, :V]
                ops.linear(t, w["dw"], w["db"], out=logits)
                acc = ops.masked_ce(logits, caption_label.reshape(-1), V, -100)
            mlm_loss = acc[0] / acc[1]
        if not cfg.ITM_task:
            return mlm_loss
        itm_logits = ops.linear_small(self.MVLBert.pool(hidden, shadow, B, S), pk["itm_w"], pk["itm_b"])
        acc = ops.masked_ce(itm_logits, image_text_label.reshape(-1), 2, -100)
        return mlm_loss.mean() + acc[0] / acc[1]


class MVLBertForImageCaption(_PackedMixin, MVLBertPretrainedModel):
    """model.py:479-550: report generation / image captioning.  `forward(image, caption, num_beams=0, learning_strategy)` is the
    teacher-forced pass (`encode_forward`, model.py:518-550: seq2seq mask, MLM head over the text positions ('unilm') or over
    [SEP], t1..t(n-1) ('normal')) -> logits [B, vocab, L] as in the reference.  `num_beams=1` is greedy decoding (`greedy_search`,
    by prefix recompute on the forward kernels); beam search (`num_beams > 1`) is outside the accelerated path."""

    def __init__(self, config, tokenizer=None):
        super().__init__(config)
        self.config = config
        assert config.is_decoder, 'config.is_decoder should be True if you want to run image caption for testing'
        self.MVLBert = MVLBert(config, add_pooling_layer=True)
        self.conv = Conv_layer(config)
        self.tokenizer = tokenizer
        self.MLM_head_seq2seq = _OnlyMLMHead(config)
        self._finish_init()

    def _pack_params(self):
        return self.MLM_head_seq2seq.parameters()

    def _pack(self):
        p, wd = self.MLM_head_seq2seq.predictions, act_dtype(self.precision)
        f32 = lambda x: x.detach().float().contiguous()
        return dict(tw=p.transform.dense.weight.detach().to(wd).contiguous(), tb=f32(p.transform.dense.bias),
                    lw=f32(p.transform.LayerNorm.weight), lb=f32(p.transform.LayerNorm.bias),
                    dw=p.decoder.weight.detach().to(wd).contiguous(), db=f32(p.decoder.bias))

    def _mlm_logits(self, rows):
        """BertOnlyMLMHead (HF modeling_bert.py:488-512) on [n, D] rows in the GEMM operand dtype -> fp32 logits [n, vocab]."""
        w = self.packed()
        t = ops.linear(rows, w["tw"], w["tb"], act=ops.ACT_GELU, out_dtype=torch.float32)
        t = ops.layernorm(t, w["lw"], w["lb"], self.config.layer_norm_eps, act_dtype(self.precision))
        V = w["dw"].shape[0]
        ld = (V + 31) // 32 * 32
        logits = torch.empty((rows.shape[0], ld), device=t.device, dtype=torch.float32)[:, :V]
        ops.linear(t, w["dw"], w["db"], out=logits)
        return logits

    def forward(self, image, caption, num_beams=0, learning_strategy="unilm", sample_mode="greedy"):
        if num_beams > 1:
            raise NotImplementedError("beam search decode (model.py:636-824) is outside the accelerated forward path; "
                                      "num_beams=0 runs the teacher-forced pass, num_beams=1 greedy decoding")
        if learning_strategy not in ("unilm", "normal"):
            raise NotImplementedError("learning_strategy:", learning_strategy, "is not implemented! Try 'unilm' or 'normal'.")
        if num_beams == 1:
            return self.greedy_search(self.conv(image), learning_strategy=learning_strategy, sample_mode=sample_mode)
        feat, hidden, shadow, B, S = self._trunk(image, caption, None, True)
        n_obj, L, D = feat.shape[1], caption.shape[1], hidden.shape[1]
        # 'unilm': hidden states of t1..tn ; 'normal': [SEP], t1..t(n-1) = the same window shifted one row up (model.py:536-544)
        first = n_obj + 2 if learning_strategy == "unilm" else n_obj + 1
        src = shadow if shadow is not None else hidden
        rows = src.view(B, S, D)[:, first:first + L].reshape(B * L, D)
        V = self.packed()["dw"].shape[0]
        return self._mlm_logits(rows).view(B, L, V).transpose(1, 2)          # batch, vocab_size, seq_len (model.py:534)

    @torch.no_grad()
    def greedy_search(self, image_feature, learning_strategy="unilm", sample_mode="greedy", max_length=None,
                      pad_token_id=None, eos_token_id=None, mask_token_id=None):
        """model.py:826-984 with `learning_strategy='unilm'`, `sample_mode='greedy'`: every step feeds the tokens generated so
        far plus one [MASK] and takes the argmax of the MLM head at the [MASK] position; finished rows emit `pad`; stops when
        every row has produced [END] or at `max_length`.  -> (input_ids [B, steps], concatenated per-step max logits).

        The reference decodes incrementally through HF's KV cache (two new tokens per step, the [MASK] entry trimmed from the
        cache afterwards, model.py:890-894).  Under the seq2seq mask every cached key/value is exactly what a forward over
        the whole prefix recomputes, so this implementation RE-RUNS the joint encoder over the growing prefix each step on the
        forward kernels (image features computed once): same arithmetic, O(L^2) instead of O(L) token-forwards."""
        if learning_strategy != "unilm" or sample_mode != "greedy":
            raise NotImplementedError("only learning_strategy='unilm' with sample_mode='greedy' decodes on the accelerated path")
        cfg = self.config
        max_length = max_length if max_length is not None else cfg.max_length
        pad = pad_token_id if pad_token_id is not None else (cfg.pad_token_id if cfg.pad_token_id is not None else 0)
        eos = eos_token_id if eos_token_id is not None else cfg.eos_token_id
        mask_id = mask_token_id if mask_token_id is not None else \
            (self.tokenizer.mask_token_id if self.tokenizer is not None else cfg.mask_token_id)
        B, dev = image_feature.shape[0], image_feature.device
        unfinished = torch.ones(B, dtype=torch.int64, device=dev)
        mask_col = torch.full((B, 1), mask_id, dtype=torch.int64, device=dev)
        input_ids, probs = None, []
        for _ in range(max_length):
            text = mask_col if input_ids is None else torch.cat([input_ids, mask_col], dim=-1)
            hidden, shadow, _, S = self.MVLBert.encode(text.contiguous(), None, image_feature, None, True)
            src = shadow if shadow is not None else hidden
            logits = self._mlm_logits(src.view(B, S, -1)[:, -1])                 # hidden state of the [MASK] position
            scores, tokens = torch.max(logits, dim=-1)
            if eos is not None:
                tokens = tokens * unfinished + pad * (1 - unfinished)
            input_ids = tokens[:, None] if input_ids is None else torch.cat([input_ids, tokens[:, None]], dim=-1)
            if eos is not None:
                unfinished = unfinished * (tokens != eos).long()
            if unfinished.max() == 0:
                break
            probs.append(scores)
        return input_ids, (torch.cat(probs, dim=-1) if probs else torch.empty(0, device=dev))
