"""CPU oracle for the MVLT multimodal forward hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32 or fp64) functional restatement of the reference algorithm, written
over a flat `state_dict` so it needs neither the reference tree nor HuggingFace module classes
and can therefore travel to the GPU box.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s baseline legs (`cpu_baseline`, `--impl reference`, the secondary `gpu_eager_baseline`, the `parity` check)
may import it; the product package (`medical_vision_langauge_transformer_b200`) never does.

Pinning status: the reference ships NO golden vectors or tests for this path (SURVEY.md §4, §8c:
"parity unpinned" on the reference side).  This restatement is therefore pinned operationally:
  * `tests/test_oracle_cpu.py::test_oracle_vs_live_reference` runs it against the real, unmodified reference modules
    (imported through `oracle/ref_shims.py`) whenever `/root/reference` is present, and
  * `oracle/make_golden.py` stores outputs OF THE REAL REFERENCE on seeded inputs/weights under
    `tests/golden/`; the other tests of `tests/test_oracle_cpu.py` check this file against them everywhere.

Every function cites the reference lines it follows.  `vfe.py` = modules/visual_feature_extractor.py,
`model.py` = modules/model.py (both under the reference root), `HF:` = transformers 5.5.0
`models/bert/modeling_bert.py` (third-party dependency, reference pin `transformers>=4.16.0`,
README.md:7; the BERT arithmetic is not vendored in the reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

SWIN_S = dict(embed_dim=96, depths=(2, 2, 18, 2), num_heads=(3, 6, 12, 24), window=7, img=224, patch=4)


# ----------------------------------------------------------------------------------------------
# Swin building blocks
# ----------------------------------------------------------------------------------------------
def relative_position_index(ws: int) -> Tensor:
    """vfe.py:203-214 — pairwise (dy+ws-1)*(2ws-1)+(dx+ws-1) over the ws*ws window tokens."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    c = torch.stack([ys.reshape(-1), xs.reshape(-1)])          # 2, N
    rel = c[:, :, None] - c[:, None, :]                        # 2, N, N   (row i minus col j)
    return (rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)


def shift_attn_mask(H: int, W: int, ws: int, shift: int, dtype=torch.float32) -> Tensor:
    """vfe.py:318-344 — region ids over the *shifted* image, -100 where ids differ. -> [nW, N, N]"""
    ids = torch.zeros(H, W, dtype=dtype)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            ids[hs, wsl] = cnt
            cnt += 1
    win = ids.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def window_partition(x: Tensor, ws: int) -> Tensor:
    """vfe.py:144-156: [B,H,W,C] -> [B*nW, ws*ws, C]"""
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def window_reverse(w: Tensor, ws: int, H: int, W: int) -> Tensor:
    """vfe.py:159-173: [B*nW, ws*ws, C] -> [B,H,W,C]"""
    C = w.shape[-1]
    B = w.shape[0] // ((H // ws) * (W // ws))
    return w.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)


def window_attention(sd: SD, p: str, xw: Tensor, heads: int, ws: int, mask: Optional[Tensor]) -> Tensor:
    """vfe.py:224-254."""
    B_, N, C = xw.shape
    hd = C // heads
    qkv = F.linear(xw, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).view(B_, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]                       # :234 scale on q
    attn = q @ k.transpose(-2, -1)
    idx = relative_position_index(ws).reshape(-1)
    bias = sd[p + "relative_position_bias_table"][idx].view(N, N, heads).permute(2, 0, 1)   # :236-238
    attn = attn + bias.unsqueeze(0)
    if mask is not None:                                                   # :241-244
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.to(attn.dtype)[None, :, None]).view(-1, heads, N, N)
    attn = attn.softmax(-1)
    out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(out, sd[p + "proj.weight"], sd[p + "proj.bias"])


def swin_block(sd: SD, p: str, x: Tensor, H: int, W: int, heads: int, ws: int, shift: int) -> Tensor:
    """vfe.py:350-387 (DropPath/Dropout are identity in eval)."""
    B, L, C = x.shape
    if min(H, W) <= ws:                                                    # :302-305
        shift, ws = 0, min(H, W)
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(B, H, W, C)
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))            # :361
    mask = shift_attn_mask(H, W, ws, shift, x.dtype) if shift > 0 else None
    a = window_attention(sd, p + "attn.", window_partition(h, ws), heads, ws, mask)
    h = window_reverse(a, ws, H, W)
    if shift > 0:
        h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))              # :378
    x = x + h.reshape(B, L, C)                                             # :384
    m = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    m = F.linear(m, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    m = F.gelu(m)                                                          # exact erf GELU, vfe.py:126
    m = F.linear(m, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + m                                                           # :385


def patch_merging(sd: SD, p: str, x: Tensor, H: int, W: int) -> Tensor:
    """vfe.py:424-445 — concat order (0,0),(1,0),(0,1),(1,1) then LN(4C) then Linear(4C,2C,no bias)."""
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).view(B, -1, 4 * C)
    x = F.layer_norm(x, (4 * C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "reduction.weight"])


def patch_embed(sd: SD, p: str, img: Tensor, patch: int = 4) -> Tensor:
    """vfe.py:557-565 — Conv2d(k=s=patch) + flatten + transpose + LN."""
    x = F.conv2d(img, sd[p + "proj.weight"], sd[p + "proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    C = x.shape[-1]
    return F.layer_norm(x, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)


def swin_forward(sd: SD, img: Tensor, prefix: str = "conv.conv.0.", cfg=SWIN_S, taps: Optional[dict] = None) -> Tensor:
    """vfe.py:676-693 (forward_features; ape=False, pos_drop identity). -> [B,49,768]"""
    x = patch_embed(sd, prefix + "patch_embed.", img, cfg["patch"])
    if taps is not None:
        taps["patch_embed"] = x
    res = cfg["img"] // cfg["patch"]
    for s, (depth, heads) in enumerate(zip(cfg["depths"], cfg["num_heads"])):
        H = W = res // (2 ** s)
        for i in range(depth):
            shift = 0 if i % 2 == 0 else cfg["window"] // 2                # vfe.py:491
            x = swin_block(sd, f"{prefix}layers.{s}.blocks.{i}.", x, H, W, heads, cfg["window"], shift)
            if taps is not None and i < 2:
                taps[f"s{s}b{i}"] = x
        if s < len(cfg["depths"]) - 1:
            x = patch_merging(sd, f"{prefix}layers.{s}.downsample.", x, H, W)
        if taps is not None:
            taps[f"stage{s}"] = x
    C = x.shape[-1]
    return F.layer_norm(x, (C,), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-5)


# ----------------------------------------------------------------------------------------------
# ResNet backbones (third-party arithmetic: torchvision/models/resnet.py, reference pin torchvision>=0.12.0, README.md:5)
# ----------------------------------------------------------------------------------------------
RESNET_LAYERS = {"resnet101": (3, 4, 23, 3), "resnet50": (3, 4, 6, 3)}


def _bn_eval(sd: SD, p: str, x: Tensor) -> Tensor:
    """nn.BatchNorm2d in eval mode (running statistics, eps 1e-5)."""
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5)


def bottleneck(sd: SD, p: str, x: Tensor, stride: int) -> Tensor:
    """torchvision resnet.py Bottleneck.forward (v1.5: stride on conv2): 1x1 -> bn -> relu -> 3x3/stride -> bn -> relu ->
    1x1 -> bn -> += identity (downsample = 1x1/stride conv + bn when present) -> relu."""
    out = F.relu(_bn_eval(sd, p + "bn1.", F.conv2d(x, sd[p + "conv1.weight"])))
    out = F.relu(_bn_eval(sd, p + "bn2.", F.conv2d(out, sd[p + "conv2.weight"], stride=stride, padding=1)))
    out = _bn_eval(sd, p + "bn3.", F.conv2d(out, sd[p + "conv3.weight"]))
    if p + "downsample.0.weight" in sd:
        x = _bn_eval(sd, p + "downsample.1.", F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride))
    return F.relu(out + x)


def resnet_forward(sd: SD, img: Tensor, prefix: str = "conv.conv.0.", layers=RESNET_LAYERS["resnet101"],
                   taps: Optional[dict] = None) -> Tensor:
    """vfe.py:14-24 `_forward_impl`: conv1 7x7/2 -> bn1 -> relu -> maxpool 3x3/2 -> layer1..4; no avgpool, no fc.
    -> [B, 2048, 7, 7]"""
    x = F.relu(_bn_eval(sd, prefix + "bn1.", F.conv2d(img, sd[prefix + "conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    if taps is not None:
        taps["stem"] = x
    for li, depth in enumerate(layers):
        for b in range(depth):
            x = bottleneck(sd, f"{prefix}layer{li + 1}.{b}.", x, 2 if (b == 0 and li > 0) else 1)
        if taps is not None:
            taps[f"layer{li + 1}"] = x
    return x


# ----------------------------------------------------------------------------------------------
# Linear-patch stem (vfe.py:47-60) and ViT-B/16 trunk (vfe.py:66-107; third-party arithmetic: torchvision
# models/vision_transformer.py + torch.nn.MultiheadAttention, reference pin torchvision>=0.12.0)
# ----------------------------------------------------------------------------------------------
def linear_patch_forward(sd: SD, img: Tensor, prefix: str = "conv.conv.0.") -> Tensor:
    """vfe.py:55-60: Conv2d(3,768,k=16,s=16) -> BatchNorm2d (eval) -> ReLU -> [B,768,14,14]."""
    x = F.conv2d(img, sd[prefix + "linear_patch.weight"], sd[prefix + "linear_patch.bias"], stride=16)
    return F.relu(_bn_eval(sd, prefix + "bn.", x))


def vit_forward(sd: SD, img: Tensor, prefix: str = "conv.conv.0.", heads: int = 12, taps: Optional[dict] = None) -> Tensor:
    """vfe.py:89-107 over torchvision's VisionTransformer: `_process_input` (conv_proj 16x16/16 -> [B,196,768]), class token,
    + pos_embedding, 12 pre-LN EncoderBlocks (ln_1 -> MultiheadAttention with packed in_proj -> + x; ln_2 -> Linear, GELU,
    Linear -> + x), final LayerNorm (eps 1e-6), and the reference drops the class token: x[:, 1:] -> [B,196,768]."""
    x = F.conv2d(img, sd[prefix + "conv_proj.weight"], sd[prefix + "conv_proj.bias"], stride=16).flatten(2).transpose(1, 2)
    B, _, D = x.shape
    x = torch.cat([sd[prefix + "class_token"].expand(B, -1, -1), x], 1) + sd[prefix + "encoder.pos_embedding"]
    n_layers = sum(1 for k in sd if k.startswith(prefix + "encoder.layers.") and k.endswith(".ln_1.weight"))
    hd = D // heads
    for i in range(n_layers):
        p = f"{prefix}encoder.layers.encoder_layer_{i}."
        h = F.layer_norm(x, (D,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-6)
        qkv = F.linear(h, sd[p + "self_attention.in_proj_weight"], sd[p + "self_attention.in_proj_bias"])
        q, k, v = (t.view(B, -1, heads, hd).transpose(1, 2) for t in qkv.chunk(3, -1))
        a = ((q @ k.transpose(-1, -2)) / math.sqrt(hd)).softmax(-1) @ v
        a = a.transpose(1, 2).reshape(B, -1, D)
        x = x + F.linear(a, sd[p + "self_attention.out_proj.weight"], sd[p + "self_attention.out_proj.bias"])
        h = F.layer_norm(x, (D,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-6)
        x = x + F.linear(F.gelu(F.linear(h, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])), sd[p + "mlp.3.weight"], sd[p + "mlp.3.bias"])
        if taps is not None and i in (0, n_layers - 1):
            taps[f"vit{i}"] = x
    x = F.layer_norm(x, (D,), sd[prefix + "encoder.ln.weight"], sd[prefix + "encoder.ln.bias"], 1e-6)
    return x[:, 1:]


def conv_layer(sd: SD, img: Tensor, taps: Optional[dict] = None) -> Tensor:
    feat = _conv_layer(sd, img, taps)
    if taps is not None:
        taps["image_feature"] = feat
    return feat


def _conv_layer(sd: SD, img: Tensor, taps: Optional[dict] = None) -> Tensor:
    """model.py:232-235,255-266 — Sequential(backbone, GELU); 4-D branch.  Swin: feature dim 768 so no resnet_fc.
    ResNet (keys `conv.conv.0.layer1...`): [B,2048,7,7] -> reshape/transpose [B,49,2048] (:258-261) -> resnet_fc (:263-264);
    the depth is read off the state_dict (layer3 has 23 blocks for resnet101, 6 for resnet50)."""
    if img.dim() == 5:                       # IU-Xray two-view input, model.py:240-253: objects of both views concatenated
        return torch.cat((_conv_layer(sd, img[:, 0], taps), _conv_layer(sd, img[:, 1])), dim=1)
    if "conv.conv.0.linear_patch.weight" in sd:   # model.py:200-201, :258-261: [B,768,14,14] -> [B,196,768]; width 768: no resnet_fc
        x = linear_patch_forward(sd, img)
        if taps is not None:
            taps["linear_patch"] = x
        x = F.gelu(x)
        return x.reshape(x.shape[0], x.shape[1], -1).transpose(1, 2)
    if "conv.conv.0.class_token" in sd:           # model.py:227-228: 3-D output, no reshape
        return F.gelu(vit_forward(sd, img, taps=taps))
    if "conv.conv.0.layer1.0.conv1.weight" in sd:
        layers = tuple(sum(1 for k in sd if k.startswith(f"conv.conv.0.layer{i}.") and k.endswith(".conv1.weight"))
                       for i in (1, 2, 3, 4))
        x = F.gelu(resnet_forward(sd, img, "conv.conv.0.", layers, taps))
        B, C = x.shape[:2]
        x = x.reshape(B, C, -1).transpose(1, 2)
        return F.linear(x, sd["conv.resnet_fc.weight"], sd["conv.resnet_fc.bias"])
    return F.gelu(swin_forward(sd, img, "conv.conv.0.", taps=taps))


# ----------------------------------------------------------------------------------------------
# Joint encoder
# ----------------------------------------------------------------------------------------------
def joint_embedding(sd: SD, text_idx: Tensor, image_feature: Tensor, cls_id=101, sep_id=102, p="MVLBert."):
    """model.py:110-160 — [CLS] img [SEP] text word-emb + type(1 for pos<=obj_end) + position; NO LayerNorm."""
    B, n_obj, _ = image_feature.shape
    obj_end = n_obj + 1
    S = text_idx.shape[1] + n_obj + 2
    we = sd[p + "word_embeddings.weight"]
    dt = image_feature.dtype
    cls = we[cls_id].to(dt).expand(B, 1, -1)
    sep = we[sep_id].to(dt).expand(B, 1, -1)
    vl = torch.cat([cls, image_feature, sep, we[text_idx].to(dt)], 1)
    pos = torch.arange(S)
    typ = (pos <= obj_end).long()                                           # :152-153
    return vl + sd[p + "token_type_embeddings.weight"][typ].to(dt) + sd[p + "position_embeddings.weight"][pos].to(dt)


def joint_attention_mask(text_idx: Tensor, n_obj: int, seq2seq: bool, image_mask: Optional[Tensor] = None,
                         dtype=torch.float32) -> Tensor:
    """model.py:118-128 + :162-183 — additive mask, 0 / -10000. [B,1,1,S] or [B,1,S,S]."""
    B, L = text_idx.shape
    S = L + n_obj + 2
    obj_end = n_obj + 1
    if seq2seq:
        r = torch.arange(S)
        m = (r[None, :] <= r[:, None]) | (r[None, :] <= obj_end)          # :121-122 (text padding ignored)
        m = m[None, None].expand(B, 1, S, S)
    else:
        ones = torch.ones(B, 1, dtype=torch.bool)
        im = torch.ones(B, n_obj, dtype=torch.bool) if image_mask is None else image_mask.bool()
        m = torch.cat([ones, im, ones, text_idx > 0], 1)[:, None, None, :]
    return (1.0 - m.to(dtype)) * -10000.0


def bert_layer(sd: SD, p: str, h: Tensor, mask: Tensor, heads: int = 12) -> Tensor:
    """HF:359-421 -> BertSelfAttention :168-207 (eager :115-140), BertSelfOutput :287-298,
    BertIntermediate :330-342, BertOutput :345-356.  Post-LN, eps 1e-12, erf GELU."""
    B, S, D = h.shape
    hd = D // heads

    def split(t):
        return t.view(B, S, heads, hd).transpose(1, 2)

    q = split(F.linear(h, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"]))
    k = split(F.linear(h, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"]))
    v = split(F.linear(h, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"]))
    s = (q @ k.transpose(2, 3)) * (hd ** -0.5) + mask
    ctx = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, S, D)
    a = F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    h = F.layer_norm(a + h, (D,), sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], 1e-12)
    f = F.gelu(F.linear(h, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    f = F.linear(f, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return F.layer_norm(f + h, (D,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)


def mvlbert(sd: SD, text_idx: Tensor, image_feature: Tensor, seq2seq: bool = False,
            image_mask: Optional[Tensor] = None, layers: int = 12, taps: Optional[dict] = None):
    """model.py:35-72 — returns (last_hidden [B,S,768], pooled [B,768]).  Pooler HF:456-468."""
    h = joint_embedding(sd, text_idx, image_feature)
    if taps is not None:
        taps["embedding"] = h
    mask = joint_attention_mask(text_idx, image_feature.shape[1], seq2seq, image_mask, h.dtype)
    for l in range(layers):
        h = bert_layer(sd, f"MVLBert.encoder.layer.{l}.", h, mask)
        if taps is not None and l in (0, layers - 1):
            taps[f"bert{l}"] = h
    pooled = torch.tanh(F.linear(h[:, 0], sd["MVLBert.pooler.dense.weight"], sd["MVLBert.pooler.dense.bias"]))
    return h, pooled


def head_transform(sd: SD, p: str, x: Tensor) -> Tensor:
    """HF:471-485 BertPredictionHeadTransform: dense + erf-GELU + LN(1e-12)."""
    x = F.gelu(F.linear(x, sd[p + "dense.weight"], sd[p + "dense.bias"]))
    return F.layer_norm(x, (x.shape[-1],), sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], 1e-12)


# ----------------------------------------------------------------------------------------------
# Task forwards
# ----------------------------------------------------------------------------------------------
def vqa_forward(sd: SD, image: Tensor, question: Tensor, taps: Optional[dict] = None):
    """model.py:329-349 -> (prob, logits) [B,result_num]; final_mlp = (Dropout, Linear) so key final_mlp.1."""
    feat = conv_layer(sd, image, taps)
    _, pooled = mvlbert(sd, question, feat, taps=taps)
    logits = F.linear(pooled, sd["final_mlp.1.weight"], sd["final_mlp.1.bias"])
    return logits.softmax(-1), logits


def retrieval_head(sd: SD, pooled: Tensor) -> Tensor:
    """model.py:434-440,464 — BertPredictionHeadTransform + Linear(768,2) -> logits [B,2]."""
    return F.linear(head_transform(sd, "final_mlp.0.", pooled), sd["final_mlp.1.weight"], sd["final_mlp.1.bias"])


def retrieval_forward(sd: SD, image: Tensor, caption: Tensor, return_logits: bool = False,
                      taps: Optional[dict] = None) -> Tensor:
    """model.py:444-476 -> softmax(dim=1) prob [B,2] (or logits when a label is supplied)."""
    feat = conv_layer(sd, image, taps)
    _, pooled = mvlbert(sd, caption, feat, taps=taps)
    logits = retrieval_head(sd, pooled)
    return logits if return_logits else logits.softmax(1)


def pretrain_forward(sd: SD, image: Tensor, caption_masked: Tensor, caption_label: Tensor,
                     image_text_label: Tensor, seq2seq: bool, mlm_task: bool = True, itm_task: bool = True,
                     return_parts: bool = False):
    """model.py:372-420.  `seq2seq` is the outcome of `random.random() < 0.5` (:390-394), which the
    caller pins.  MLM head HF:488-512 on hidden[:, 51:51+L]; CE ignore_index=-100, mean over
    labelled positions; ITM Linear(768,2) + CE."""
    feat = conv_layer(sd, image)
    h, pooled = mvlbert(sd, caption_masked, feat, seq2seq=seq2seq)
    n_obj = feat.shape[1]
    text = h[:, n_obj + 2: n_obj + 2 + caption_masked.shape[1]]
    p = "MLM_head_seq2seq.predictions." if seq2seq else "MLM_head_bidir.predictions."
    t = head_transform(sd, p + "transform.", text)
    logits = F.linear(t, sd[p + "decoder.weight"], sd[p + "decoder.bias"])
    mlm = F.cross_entropy(logits.transpose(1, 2), caption_label, ignore_index=-100) if mlm_task else torch.zeros(1, 1)
    if not itm_task:
        return mlm
    itm_logits = F.linear(pooled, sd["ITM_mlp.weight"], sd["ITM_mlp.bias"])
    itm = F.cross_entropy(itm_logits, image_text_label)
    if return_parts:
        return mlm, itm, logits, itm_logits
    return mlm.mean() + itm.mean()


def caption_encode_forward(sd: SD, image: Tensor, caption: Tensor, learning_strategy: str = "unilm") -> Tensor:
    """model.py:518-550 (`MVLBertForImageCaption.encode_forward`, reached with num_beams=0): seq2seq mask, MLM head (HF:488-512)
    over t1..tn ('unilm') or [SEP], t1..t(n-1) ('normal') -> logits [B, vocab, L]."""
    feat = conv_layer(sd, image)
    h, _ = mvlbert(sd, caption, feat, seq2seq=True)
    n_obj, L = feat.shape[1], caption.shape[1]
    first = n_obj + 2 if learning_strategy == "unilm" else n_obj + 1
    p = "MLM_head_seq2seq.predictions."
    t = head_transform(sd, p + "transform.", h[:, first:first + L])
    return F.linear(t, sd[p + "decoder.weight"], sd[p + "decoder.bias"]).transpose(1, 2)


def caption_greedy_decode(sd: SD, image: Tensor, max_length: int, mask_id: int = 103, eos_id: int = 104, pad_id: int = 0):
    """model.py:826-984 (`greedy_search`, learning_strategy='unilm', sample_mode='greedy') restated WITHOUT the KV cache: each
    step runs the joint encoder over [t1..tk, [MASK]] with the seq2seq mask and takes the argmax of the MLM head at the [MASK]
    position (:881-897); finished rows emit pad (:935); the [MASK] entry never enters the next step's context (:890-894 trims
    it from the cache), which is exactly what re-encoding the prefix does.  NOTE: the reference's own decode loop does not run
    under the installed transformers 5.5 (BertEncoder no longer returns tuple caches), so this function is pinned only by
    its agreement with `caption_encode_forward` (teacher-forced logits at the same positions) — "parity unpinned" for decode.
    -> (input_ids [B, steps], per-step max logits as the reference concatenates them, list of per-step top-2 margins)."""
    feat = conv_layer(sd, image)
    B = feat.shape[0]
    p = "MLM_head_seq2seq.predictions."
    unfinished = torch.ones(B, dtype=torch.int64)
    mask_col = torch.full((B, 1), mask_id, dtype=torch.int64)
    input_ids, probs, margins = None, [], []
    for _ in range(max_length):
        text = mask_col if input_ids is None else torch.cat([input_ids, mask_col], -1)
        h, _ = mvlbert(sd, text, feat, seq2seq=True)
        logits = F.linear(head_transform(sd, p + "transform.", h[:, -1]), sd[p + "decoder.weight"], sd[p + "decoder.bias"])
        top2 = logits.topk(2, -1).values
        margins.append(top2[:, 0] - top2[:, 1])
        scores, tokens = logits.max(-1)
        tokens = tokens * unfinished + pad_id * (1 - unfinished)
        input_ids = tokens[:, None] if input_ids is None else torch.cat([input_ids, tokens[:, None]], -1)
        unfinished = unfinished * (tokens != eos_id).long()
        if unfinished.max() == 0:
            break
        probs.append(scores)
    return input_ids, (torch.cat(probs, -1) if probs else torch.empty(0)), margins


# ----------------------------------------------------------------------------------------------
# N x N retrieval scoring and ranking
# ----------------------------------------------------------------------------------------------
def retrieval_score_matrix(sd: SD, images: Tensor, captions: Tensor, chunk: int = 64) -> Tensor:
    """run_retrieval.py:133-145,198-213 — score every (image i, caption j): prob[:,1] of the pair, row-major
    [N_img, N_cap].  The reference recomputes the Swin trunk per pair; Conv_layer depends only on the image,
    so it is evaluated once per image here (identical arithmetic per pair)."""
    feats = torch.cat([conv_layer(sd, images[i:i + chunk]) for i in range(0, images.shape[0], chunk)])
    N_i, N_c = feats.shape[0], captions.shape[0]
    out = torch.empty(N_i, N_c, dtype=feats.dtype)
    for i in range(N_i):
        for j0 in range(0, N_c, chunk):
            cap = captions[j0:j0 + chunk]
            _, pooled = mvlbert(sd, cap, feats[i:i + 1].expand(cap.shape[0], -1, -1))
            out[i, j0:j0 + chunk] = retrieval_head(sd, pooled).softmax(1)[:, 1]
    return out


def compute_ranks(scores, labels):
    """run_retrieval.py:220-249 — per row: argsort descending (numpy argsort reversed), index of the first
    label==1 (N if none); then the same per column.  numpy on purpose: tie order must match the reference."""
    import numpy as np
    scores, labels = np.asarray(scores), np.asarray(labels)
    N = scores.shape[1]

    def one(sim, lab):
        ranks = []
        for l, s in zip(lab, sim):
            inds = np.argsort(s)[::-1]
            hit = np.nonzero(l[inds] == 1)[0]
            ranks.append(int(hit[0]) if hit.size else N)
        return ranks

    return one(scores, labels), one(scores.T, labels.T)


def recall_at(ranks, ks=(1, 5, 10)):
    """run_retrieval.py:283-294."""
    return [sum(r < k for r in ranks) / len(ranks) for k in ks]


def resnet_flops(layers=RESNET_LAYERS["resnet101"], img: int = 224) -> float:
    """Algorithmic FLOPs (MAC*2) of the Bottleneck trunk (stem + layer1-4, no fc) per image — SURVEY.md §8d: 15.60 G for
    ResNet-101 at 224^2."""
    H = img // 2
    f = 2.0 * H * H * 64 * 3 * 49
    H //= 2
    inpl = 64
    for li, depth in enumerate(layers):
        planes = 64 * 2 ** li
        for b in range(depth):
            stride = 2 if (b == 0 and li > 0) else 1
            Ho = H // stride
            f += 2.0 * H * H * inpl * planes + 2.0 * Ho * Ho * planes * planes * 9 + 2.0 * Ho * Ho * planes * planes * 4
            if b == 0:
                f += 2.0 * Ho * Ho * inpl * planes * 4
            inpl, H = planes * 4, Ho
    return f


def flops_per_pair(L: int = 80, conv: str = "swintransformer") -> float:
    """Algorithmic FLOPs (MAC*2, unpadded) of one Swin-S + BERT-base pair — BASELINE.md §3 (conv="resnet101"/"resnet50":
    the Bottleneck trunk + resnet_fc instead of Swin-S)."""
    if conv in ("linear", "vit"):              # 196 image tokens: S = 198 + L
        S = 198 + L
        bert = 12 * (2 * S * 768 * (3 * 768 + 768 + 2 * 3072) + 2 * 2 * 12 * S * S * 64)
        trunk = 2 * 196 * 768 * 768
        if conv == "vit":
            trunk += 12 * (2 * 197 * 768 * (3 * 768 + 768 + 2 * 3072) + 2 * 2 * 12 * 197 * 197 * 64)
        return trunk + bert + 2 * 768 * 768 * 2
    if conv in RESNET_LAYERS:
        S = 51 + L
        bert = 12 * (2 * S * 768 * (3 * 768 + 768 + 2 * 3072) + 2 * 2 * 12 * S * S * 64)
        return resnet_flops(RESNET_LAYERS[conv]) + 2 * 49 * 2048 * 768 + bert + 2 * 768 * 768 * 2
    swin = 0.0
    res = 56
    swin += 2 * 3136 * 48 * 96
    for s, (d, nh) in enumerate(zip(SWIN_S["depths"], SWIN_S["num_heads"])):
        T, C = (res // 2 ** s) ** 2, 96 * 2 ** s
        per_block = 2 * T * C * (3 * C + C + 8 * C) + 2 * 2 * (T // 49) * nh * 49 * 49 * 32
        swin += d * per_block
        if s < 3:
            swin += 2 * (T // 4) * 4 * C * 2 * C
    S = 51 + L
    bert = 12 * (2 * S * 768 * (3 * 768 + 768 + 2 * 3072) + 2 * 2 * 12 * S * S * 64)
    head = 2 * 768 * 768 * 2
    return swin + bert + head
