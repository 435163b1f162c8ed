"""First slice of the training step: forward-with-saved-activations and backward of ONE BertLayer
(HF modeling_bert.py:359-421; `loss.backward()` of run_pretrain.py:177-184 / run_vqa.py:108-112 restricted to one encoder layer).

What is here: every dgrad (dX = dY . W) and wgrad (dW = dY^T . X) of the layer's four nn.Linear sites runs on the tcgen05 GEMM
(`ops.linear`, C = A . B^T) over transposed bf16 operands; LayerNorm / GELU / attention backward, bias gradients and the
transposes are CUDA kernels of csrc/backward.cu.  bf16 operands, fp32 accumulation, fp32 residual-stream gradients — the mixed
precision of the forward.  Parity: tests/test_backward_gpu.py against torch.autograd of the oracle's `bert_layer`.

`bert_encoder_forward / bert_encoder_backward` chain the layers (HF modeling_bert.py:424-453 BertEncoder) and return the gradients under
the reference's `encoder.layer.{l}.…` keys.

What is NOT here: the embedding / pooler / head backward, the Swin backward, dropout / DropPath in train mode, the optimizer, the
gradient all-reduce, and tensor-core versions of the attention / row backward kernels.  The forward below uses the unfused kernel chain
because it has to keep the pre-LayerNorm sums and the pre-GELU activation that the fused inference kernels never write."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops

LAYER_PARAM_KEYS = (
    "attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight", "attention.self.key.bias",
    "attention.self.value.weight", "attention.self.value.bias", "attention.output.dense.weight", "attention.output.dense.bias",
    "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
    "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias")


def pack_layer(sd: Dict[str, torch.Tensor], prefix: str, device="cuda") -> Dict[str, torch.Tensor]:
    """Kernel-layout copies of one BertLayer's parameters (reference keys under `prefix`): bf16 [N, K] weights for the forward GEMMs
    (Q|K|V concatenated), their transposes [K, N] for the dgrad GEMMs, fp32 biases and LayerNorm parameters."""
    g = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32).contiguous()
    w = {"qkv_w": torch.cat([g("attention.self.query.weight"), g("attention.self.key.weight"), g("attention.self.value.weight")]).bfloat16(),
         "qkv_b": torch.cat([g("attention.self.query.bias"), g("attention.self.key.bias"), g("attention.self.value.bias")]),
         "ao_w": g("attention.output.dense.weight").bfloat16(), "ao_b": g("attention.output.dense.bias"),
         "ln1_w": g("attention.output.LayerNorm.weight"), "ln1_b": g("attention.output.LayerNorm.bias"),
         "fi_w": g("intermediate.dense.weight").bfloat16(), "fi_b": g("intermediate.dense.bias"),
         "fo_w": g("output.dense.weight").bfloat16(), "fo_b": g("output.dense.bias"),
         "ln2_w": g("output.LayerNorm.weight"), "ln2_b": g("output.LayerNorm.bias")}
    for k in ("qkv", "ao", "fi", "fo"):
        w[k + "_t"] = ops.transpose_to_bf16(w[k + "_w"], pad_to=8)          # [K, N]: the B operand of dX = dY . W
    return w


def bert_layer_forward(w, h: torch.Tensor, kmask: Optional[torch.Tensor], B: int, S: int, heads: int = 12, seq2seq: bool = False,
                       obj_end: int = 50, eps: float = 1e-12) -> Tuple[torch.Tensor, dict]:
    """One post-LN BertLayer on h (fp32 [B*S, D]) -> (out fp32 [B*S, D], saved activations for `bert_layer_backward`)."""
    hb = h.bfloat16()
    qkv = ops.linear(hb, w["qkv_w"], w["qkv_b"])
    if kmask is None:
        kmask = torch.zeros(B, S, device=h.device, dtype=torch.float32)
    ctx = ops.joint_attention(qkv, kmask, B, S, heads, seq2seq, obj_end)
    s1 = ops.linear(ctx, w["ao_w"], w["ao_b"], residual=h, out_dtype=torch.float32)            # HF:295-297 before the LayerNorm
    h1, h1b = ops.layernorm(s1, w["ln1_w"], w["ln1_b"], eps, torch.float32, bf16_copy=True)
    u = ops.linear(h1b, w["fi_w"], w["fi_b"])                                                  # pre-activation (kept for gelu')
    f = ops.linear(h1b, w["fi_w"], w["fi_b"], act=ops.ACT_GELU)                                # HF:338-341
    s2 = ops.linear(f, w["fo_w"], w["fo_b"], residual=h1, out_dtype=torch.float32)             # HF:352-354 before the LayerNorm
    out = ops.layernorm(s2, w["ln2_w"], w["ln2_b"], eps, torch.float32)
    saved = dict(hb=hb, qkv=qkv, kmask=kmask, ctx=ctx, s1=s1, h1b=h1b, u=u, f=f, s2=s2, B=B, S=S, heads=heads, seq2seq=seq2seq,
                 obj_end=obj_end, eps=eps)
    return out, saved


def _wgrad(dy_b: torch.Tensor, x_b: torch.Tensor) -> torch.Tensor:
    """dW [N, K] = dY^T [N, M] . X [M, K] as C = A . B^T with A = dY^T, B = X^T (bf16, rows padded with zeros), fp32 out."""
    return ops.linear(ops.transpose_to_bf16(dy_b), ops.transpose_to_bf16(x_b), out_dtype=torch.float32)


def bert_layer_backward(w, saved: dict, dout: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """-> (dh fp32 [B*S, D], {reference parameter key (relative to the layer prefix): gradient fp32})."""
    sv = saved
    D = dout.shape[1]
    # output: LayerNorm(s2), s2 = h1 + f . Wf^T + bf
    ds2, ds2b, dg2, db2 = ops.layernorm_bwd(dout.contiguous(), sv["s2"], w["ln2_w"], sv["eps"])
    g = {"output.LayerNorm.weight": dg2, "output.LayerNorm.bias": db2, "output.dense.bias": ops.colsum(ds2),
         "output.dense.weight": _wgrad(ds2b, sv["f"])}
    df = ops.linear(ds2b, w["fo_t"])                                                           # [M, 3072] bf16
    # intermediate: f = gelu(u), u = h1 . Wi^T + bi
    du = ops.gelu_bwd(sv["u"], df)
    g["intermediate.dense.bias"] = ops.colsum(du)
    g["intermediate.dense.weight"] = _wgrad(du, sv["h1b"])
    dh1 = ops.linear(du, w["fi_t"], residual=ds2, out_dtype=torch.float32)                     # + the residual branch of s2
    # attention output: h1 = LayerNorm(s1), s1 = h + ctx . Wo^T + bo
    ds1, ds1b, dg1, db1 = ops.layernorm_bwd(dh1, sv["s1"], w["ln1_w"], sv["eps"])
    g["attention.output.LayerNorm.weight"], g["attention.output.LayerNorm.bias"] = dg1, db1
    g["attention.output.dense.bias"] = ops.colsum(ds1)
    g["attention.output.dense.weight"] = _wgrad(ds1b, sv["ctx"])
    dctx = ops.linear(ds1b, w["ao_t"])
    # self-attention core and the packed Q|K|V projection
    dqkv = ops.joint_attention_bwd(sv["qkv"], None if sv["seq2seq"] else sv["kmask"], dctx, sv["B"], sv["S"], sv["heads"], sv["seq2seq"],
                                   sv["obj_end"])
    dbqkv, dwqkv = ops.colsum(dqkv), _wgrad(dqkv, sv["hb"])
    for i, name in enumerate(("query", "key", "value")):
        g[f"attention.self.{name}.weight"] = dwqkv[i * D:(i + 1) * D]
        g[f"attention.self.{name}.bias"] = dbqkv[i * D:(i + 1) * D]
    dh = ops.linear(dqkv, w["qkv_t"], residual=ds1, out_dtype=torch.float32)                   # + the residual branch of s1
    return dh, g


def pack_encoder(sd: Dict[str, torch.Tensor], prefix: str = "MVLBert.encoder.layer.", n_layers: int = 12, device="cuda"):
    """Kernel-layout copies of every encoder layer (reference keys `{prefix}{l}.…`)."""
    return [pack_layer(sd, f"{prefix}{l}.", device) for l in range(n_layers)]


def bert_encoder_forward(ws, h: torch.Tensor, kmask: Optional[torch.Tensor], B: int, S: int, heads: int = 12, seq2seq: bool = False,
                         obj_end: int = 50, eps: float = 1e-12):
    """The post-LN layer stack on h (fp32 [B*S, D]) -> (last hidden state fp32, per-layer saved activations)."""
    saved = []
    for w in ws:
        h, sv = bert_layer_forward(w, h, kmask, B, S, heads, seq2seq, obj_end, eps)
        saved.append(sv)
    return h, saved


def bert_encoder_backward(ws, saved, dout: torch.Tensor):
    """-> (dh of the encoder input, {`{l}.{parameter key}`: gradient}) — layers walked in reverse, each layer's saved activations
    released as soon as its gradients are formed."""
    grads: Dict[str, torch.Tensor] = {}
    d = dout
    for l in range(len(ws) - 1, -1, -1):
        d, g = bert_layer_backward(ws[l], saved[l], d)
        saved[l] = None
        for k, v in g.items():
            grads[f"{l}.{k}"] = v
    return d, grads
