"""SURVEY §8 f-2, first slice: backward of ONE BertLayer through the tcgen05 GEMM + csrc/backward.cu against torch.autograd of the
oracle's `bert_layer` (the CPU restatement of HF modeling_bert.py:359-421 that the forward parity is pinned to)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from medical_vision_langauge_transformer_b200 import _lib
    _lib.ensure_init()
    return torch.device("cuda")


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def test_transpose_colsum_gelu_bwd(cuda):
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(333, 200, seed=1).cuda()
    t = ops.transpose_to_bf16(x)
    assert t.shape == (200, 384) and torch.equal(t[:, :333], x.bfloat16().t()) and (t[:, 333:] == 0).all()
    xb = x.bfloat16()
    assert torch.equal(ops.transpose_to_bf16(xb, pad_to=8)[:, :333], xb.t())
    assert relerr(ops.colsum(x), x.double().sum(0)) < 1e-6 and relerr(ops.colsum(xb), xb.double().sum(0)) < 1e-6
    assert torch.equal(ops.colsum(x), ops.colsum(x))
    u = rnd(64, 3072, seed=2, scale=2.0).cuda().bfloat16(); df = rnd(64, 3072, seed=3).cuda().bfloat16()
    uu = u.float().requires_grad_(True)
    F.gelu(uu).backward(df.float())
    assert relerr(ops.gelu_bwd(u, df), uu.grad) < 5e-3                       # one bf16 rounding of the product


@pytest.mark.parametrize("rows,C", [(1, 768), (77, 768), (4192, 768), (300, 128), (50, 1024)])
def test_layernorm_bwd(cuda, rows, C):
    from medical_vision_langauge_transformer_b200 import ops
    x = (rnd(rows, C, seed=1, scale=2.0) + 0.3).cuda(); dy = rnd(rows, C, seed=2).cuda()
    g, b = (1 + rnd(C, seed=3, scale=0.1)).cuda(), rnd(C, seed=4, scale=0.1).cuda()
    xr, gr, br = x.double().requires_grad_(True), g.double().requires_grad_(True), b.double().requires_grad_(True)
    F.layer_norm(xr, (C,), gr, br, 1e-12).backward(dy.double())
    dx, dxb, dg, db = ops.layernorm_bwd(dy, x, g, 1e-12)
    assert relerr(dx, xr.grad) < 1e-5 and relerr(dg, gr.grad) < 1e-5 and relerr(db, br.grad) < 1e-5
    assert torch.equal(dxb, dx.bfloat16())
    again = ops.layernorm_bwd(dy, x, g, 1e-12)
    assert torch.equal(again[2], dg) and torch.equal(again[3], db)          # fixed reduction order


@pytest.mark.parametrize("B,S,seq2seq", [(2, 131, False), (3, 74, True), (1, 160, False), (2, 33, True)])
def test_joint_attention_bwd(cuda, B, S, seq2seq):
    from medical_vision_langauge_transformer_b200 import ops
    heads, D = 12, 768
    qkv = rnd(B * S, 3 * D, seed=5).cuda().bfloat16(); dctx = rnd(B * S, D, seed=6).cuda().bfloat16()
    kmask = torch.zeros(B, S); kmask[:, S - 7:] = -10000.0
    obj_end = 20
    q, k, v = (t.float().view(B, S, heads, 64).transpose(1, 2).requires_grad_(True) for t in qkv.cpu().split(D, dim=1))
    if seq2seq:
        r = torch.arange(S)
        m = torch.where((r[None, :] <= r[:, None]) | (r[None, :] <= obj_end), 0.0, -10000.0)[None, None]
    else:
        m = kmask[:, None, None, :]
    ctx = ((q @ k.transpose(2, 3)) / 8 + m).softmax(-1) @ v
    ctx.transpose(1, 2).reshape(B * S, D).backward(dctx.float().cpu())
    ref = torch.cat([t.grad.transpose(1, 2).reshape(B * S, D) for t in (q, k, v)], 1)
    got = ops.joint_attention_bwd(qkv, None if seq2seq else kmask.cuda(), dctx, B, S, heads, seq2seq, obj_end)
    for i, name in enumerate("qkv"):
        e = relerr(got[:, i * D:(i + 1) * D], ref[:, i * D:(i + 1) * D])
        assert e < 1.5e-2, (name, e)                                          # bf16 P / dS and bf16 outputs
    assert torch.equal(got, ops.joint_attention_bwd(qkv, None if seq2seq else kmask.cuda(), dctx, B, S, heads, seq2seq, obj_end))


@pytest.mark.parametrize("B,L,seq2seq", [(4, 80, False), (2, 23, True), (64, 80, False)])
def test_bert_layer_backward_vs_oracle_autograd(cuda, B, L, seq2seq):
    """dh and the 16 parameter gradients of one encoder layer (synthetic 'stress' weights of layer 0) against fp32 autograd of the
    oracle on the host: bf16 operands / fp32 accumulation — bar 2e-2 of each gradient's max-abs (measured ~5e-3)."""
    from medical_vision_langauge_transformer_b200 import synth, training
    from medical_vision_langauge_transformer_b200.modules import config as C, model as M
    from oracle import mvlt_oracle as O
    model = M.MVLBertForVQA(C.offline_config("vqa", max_length=L)).eval()
    sd = {k: v.clone() for k, v in synth.load_synth(model, 0, "stress").items()}
    prefix = "MVLBert.encoder.layer.0."
    n_obj, D = 49, 768
    S = n_obj + 2 + L
    ids = synth.synth_token_ids(B, L, 3)
    h = rnd(B, S, D, seed=11)
    dout = rnd(B, S, D, seed=12, scale=0.1)
    mask = O.joint_attention_mask(ids, n_obj, seq2seq)
    params = {k: sd[prefix + k].clone().requires_grad_(True) for k in training.LAYER_PARAM_KEYS}
    href = h.clone().requires_grad_(True)
    torch.set_num_threads(os.cpu_count() or 1)
    out_ref = O.bert_layer({prefix + k: v for k, v in params.items()}, prefix, href, mask)
    out_ref.backward(dout)
    w = training.pack_layer(sd, prefix)
    kmask = None if seq2seq else mask.view(B, S).cuda().contiguous()
    out, saved = training.bert_layer_forward(w, h.view(B * S, D).cuda(), kmask, B, S, 12, seq2seq, n_obj + 1)
    assert relerr(out, out_ref.detach().view(B * S, D)) < 1e-2
    dh, grads = training.bert_layer_backward(w, saved, dout.view(B * S, D).cuda())
    worst = {"dh": relerr(dh, href.grad.view(B * S, D))}
    assert set(grads) == set(training.LAYER_PARAM_KEYS)
    for k in training.LAYER_PARAM_KEYS:
        assert grads[k].shape == params[k].shape, k
        worst[k] = relerr(grads[k], params[k].grad)
    # the key bias adds the same q . b_k to every score of a row and cancels in the softmax: its true gradient is 0 (autograd
    # returns ~1e-9 of rounding noise), so the error is measured against the scale of the query-bias gradient instead
    kb = "attention.self.key.bias"
    worst[kb] = ((grads[kb].cpu() - params[kb].grad).abs().max() / params["attention.self.query.bias"].grad.abs().max()).item()
    print(f"B={B} L={L} seq2seq={seq2seq}: worst relerr {max(worst.values()):.2e} ({max(worst, key=worst.get)})")
    assert max(worst.values()) < 2e-2, worst
    dh2, grads2 = training.bert_layer_backward(w, saved, dout.view(B * S, D).cuda())
    assert torch.equal(dh, dh2) and all(torch.equal(grads[k], grads2[k]) for k in grads)      # deterministic


@pytest.mark.parametrize("n_layers,B,L", [(3, 2, 80), (12, 2, 23)])
def test_bert_encoder_backward_vs_oracle_autograd(cuda, n_layers, B, L):
    """The layer stack chained (HF:424-453): dh of the encoder input and every layer's 16 parameter gradients against fp32 autograd of
    the oracle's layers on the host.  Errors of the bf16 backward compound through the depth: bar 3e-2 of each gradient's max-abs."""
    from medical_vision_langauge_transformer_b200 import synth, training
    from medical_vision_langauge_transformer_b200.modules import config as C, model as M
    from oracle import mvlt_oracle as O
    model = M.MVLBertForVQA(C.offline_config("vqa", max_length=L)).eval()
    sd = {k: v.clone() for k, v in synth.load_synth(model, 0, "stress").items()}
    n_obj, D = 49, 768
    S = n_obj + 2 + L
    ids = synth.synth_token_ids(B, L, 5)
    h = rnd(B, S, D, seed=21)
    dout = rnd(B, S, D, seed=22, scale=0.1)
    mask = O.joint_attention_mask(ids, n_obj, False)
    prefix = "MVLBert.encoder.layer."
    params = {f"{l}.{k}": sd[f"{prefix}{l}.{k}"].clone().requires_grad_(True) for l in range(n_layers) for k in training.LAYER_PARAM_KEYS}
    href = h.clone().requires_grad_(True)
    torch.set_num_threads(os.cpu_count() or 1)
    x = href
    for l in range(n_layers):
        x = O.bert_layer({f"{prefix}{l}.{k}": params[f"{l}.{k}"] for k in training.LAYER_PARAM_KEYS}, f"{prefix}{l}.", x, mask)
    x.backward(dout)
    ws = training.pack_encoder(sd, prefix, n_layers)
    out, saved = training.bert_encoder_forward(ws, h.view(B * S, D).cuda(), mask.view(B, S).cuda().contiguous(), B, S, 12, False, n_obj + 1)
    assert relerr(out, x.detach().view(B * S, D)) < 1e-2
    dh, grads = training.bert_encoder_backward(ws, saved, dout.view(B * S, D).cuda())
    assert set(grads) == set(params) and all(sv is None for sv in saved)
    worst = {"dh": relerr(dh, href.grad.view(B * S, D))}
    for k in params:
        if k.endswith("attention.self.key.bias"):          # identically zero in exact arithmetic: measured on the query-bias scale
            qb = params[k.replace("key", "query")].grad.abs().max()
            worst[k] = ((grads[k].cpu() - params[k].grad).abs().max() / qb).item()
        else:
            worst[k] = relerr(grads[k], params[k].grad)
    print(f"{n_layers} layers, B={B}, L={L}: worst relerr {max(worst.values()):.2e} ({max(worst, key=worst.get)})")
    assert max(worst.values()) < 3e-2, {k: v for k, v in worst.items() if v > 1e-2}
