"""Per-kernel parity on the GPU: every C-ABI entry point against a plain torch fp32 restatement of the same op
(the oracle's functions where one exists).  Tolerances: fp32 kernels 1e-4 relative to the output scale,
bf16 kernels 1e-2 (BASELINE.json north_star)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def rnd(*shape, seed=0, scale=1.0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device)


GEMM_SHAPES = [
    # (M, N, K)  — the real call sites at small batch + ragged edges
    (128, 128, 64), (256, 256, 128), (3136 * 2, 288, 96), (3136 * 2, 96, 96), (3136 * 2, 384, 96), (3136 * 2, 96, 384),
    (784 * 2, 576, 192), (784 * 2, 192, 768), (196 * 2, 1152, 384), (196 * 2, 384, 1536), (98, 2304, 768),
    (98, 768, 3072), (262, 2304, 768), (262, 3072, 768), (262, 768, 3072), (1568, 192, 384), (64, 224, 768),
    (3, 768, 768), (160, 30522, 768), (1000, 160, 80),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tc_plain(cuda, M, N, K):
    from medical_vision_langauge_transformer_b200 import ops
    a = rnd(M, K, seed=1).bfloat16()
    w = rnd(N, K, seed=2, scale=1 / math.sqrt(K)).bfloat16()
    ldc = (N + 31) // 32 * 32
    out_full = torch.zeros(M, ldc, device="cuda", dtype=torch.float32)
    out = out_full[:, :N]
    ops.linear(a, w, out=out)
    ref = a.float() @ w.float().t()
    assert relerr(out, ref) < 2e-3, (M, N, K)
    assert out_full[:, N:].abs().max().item() == 0 if ldc > N else True


@pytest.mark.parametrize("block_n", [32, 64, 96, 128, 192, 256])
def test_gemm_tc_block_n(cuda, block_n):
    from medical_vision_langauge_transformer_b200 import ops
    M, N, K = 777, 1152, 384
    a = rnd(M, K, seed=3).bfloat16()
    w = rnd(N, K, seed=4, scale=1 / math.sqrt(K)).bfloat16()
    out = ops.linear(a, w, out_dtype=torch.float32, block_n=block_n)
    assert relerr(out, a.float() @ w.float().t()) < 2e-3


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("res", [None, torch.float32])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_tc_epilogues(cuda, act, res, out_dtype):
    from medical_vision_langauge_transformer_b200 import _lib, ops
    M, N, K = 1000, 768, 384
    a = rnd(M, K, seed=5).bfloat16()
    w = rnd(N, K, seed=6, scale=1 / math.sqrt(K)).bfloat16()
    bias = rnd(N, seed=7)
    r = None if res is None else rnd(M, N, seed=8).to(res)
    if r is not None and out_dtype == torch.bfloat16:
        # the fused residual is the model's fp32 residual stream: fp32 in, fp32 out; anything else is rejected, not emulated
        with pytest.raises(_lib.MvltNativeError):
            ops.linear(a, w, bias, act=act, residual=r, out_dtype=out_dtype)
        return
    out = ops.linear(a, w, bias, act=act, residual=r, out_dtype=out_dtype)
    ref = a.float() @ w.float().t() + bias
    ref = F.gelu(ref) if act == 1 else torch.tanh(ref) if act == 2 else ref
    if r is not None:
        ref = ref + r.float()
    assert relerr(out, ref) < (1e-2 if out_dtype == torch.bfloat16 else 2e-3)


@pytest.mark.parametrize("M,N,K,bn", [(300, 96, 96, 0), (300, 96, 96, 32), (517, 192, 768, 64), (262, 768, 3072, 0),
                                      (1000, 288, 96, 192), (130, 1000, 64, 0), (6272, 384, 1536, 128)])
def test_gemm_tc_residual_ragged_and_tile_widths(cuda, M, N, K, bn):
    """fp32 residual epilogue (TMA-prefetched chunks, in place and out of place) at ragged M/N edges and with tile
    widths whose chunk count does not divide evenly among the epilogue warps."""
    from medical_vision_langauge_transformer_b200 import ops
    a = rnd(M, K, seed=21).bfloat16()
    w = rnd(N, K, seed=22, scale=1 / math.sqrt(K)).bfloat16()
    bias = rnd(N, seed=23)
    x = rnd(M, N, seed=24)
    ref = x + a.float() @ w.float().t() + bias
    out = ops.linear(a, w, bias, residual=x, out_dtype=torch.float32, block_n=bn)      # out of place (BERT)
    assert relerr(out, ref) < 2e-3
    ops.linear(a, w, bias, residual=x, out=x, block_n=bn)                               # in place (Swin)
    assert relerr(x, ref) < 2e-3
    g = ops.linear(a, w, bias, act=1, out_dtype=torch.bfloat16, block_n=bn)             # bf16 + GELU, same shapes
    assert relerr(g, F.gelu(a.float() @ w.float().t() + bias)) < 1e-2


def test_gemm_tc_gelu_accuracy(cuda):
    """The packed-math erf-GELU epilogue against torch's exact erf GELU in fp32 output: abs error ~1e-6 + GEMM rounding."""
    from medical_vision_langauge_transformer_b200 import ops
    M, N, K = 512, 256, 64
    a = (rnd(M, K, seed=31) * 3).bfloat16()
    w = rnd(N, K, seed=32, scale=0.25).bfloat16()
    out = ops.linear(a, w, act=1, out_dtype=torch.float32)
    pre = a.float() @ w.float().t()
    assert (out - F.gelu(pre)).abs().max().item() < 2e-5 * max(1.0, pre.abs().max().item())


def test_gemm_tc_inplace_residual_and_strided_a(cuda):
    from medical_vision_langauge_transformer_b200 import ops
    # residual == out (Swin proj/fc2), and A = hidden[:, 0] with row stride S*D (pooler)
    M, N, K = 512, 384, 384
    a = rnd(M, K, seed=9).bfloat16()
    w = rnd(N, K, seed=10, scale=0.05).bfloat16()
    x = rnd(M, N, seed=11)
    ref = x + a.float() @ w.float().t()
    ops.linear(a, w, residual=x, out=x)
    assert relerr(x, ref) < 2e-3
    h = rnd(7, 131, 768, seed=12).bfloat16()
    wp = rnd(768, 768, seed=13, scale=0.03).bfloat16()
    out = ops.linear(h[:, 0], wp, rnd(768, seed=14), act=2, out_dtype=torch.float32)
    ref = torch.tanh(h[:, 0].float() @ wp.float().t() + rnd(768, seed=14))
    assert relerr(out, ref) < 2e-3


@pytest.mark.parametrize("C,M", [(96, 6272), (96, 300), (96, 1), (96, 50000), (192, 1568), (192, 129), (192, 40001),
                                 (384, 392), (384, 12544), (384, 127)])
def test_swin_mlp_fused(cuda, C, M):
    """x += fc2(gelu(fc1(LN(x)))) as one kernel (vfe.py:385 + :136-139) against (a) a torch restatement that rounds the
    LN output and the hidden activation to bf16 where the kernel does and (b) the unfused kernel chain, in place, at
    full tiles, ragged last tiles, a single row, and (C = 96 / 192: persistent kernel) more tiles than SMs."""
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(M, C, seed=C + M) * 2 + 0.3
    g, b = 1 + rnd(C, seed=1, scale=0.1), rnd(C, seed=2, scale=0.1)
    w1, b1 = rnd(4 * C, C, seed=3, scale=C ** -0.5).bfloat16(), rnd(4 * C, seed=4, scale=0.2)
    w2, b2 = rnd(C, 4 * C, seed=5, scale=(4 * C) ** -0.5).bfloat16(), rnd(C, seed=6, scale=0.2)
    a_ref = F.layer_norm(x, (C,), g, b, 1e-5).bfloat16().float()
    h_ref = F.gelu(a_ref @ w1.float().t() + b1).bfloat16().float()
    ref = x + h_ref @ w2.float().t() + b2
    a = ops.layernorm(x, g, b, 1e-5, torch.bfloat16)
    h = ops.linear(a, w1, b1, act=ops.ACT_GELU)
    chain = ops.linear(h, w2, b2, residual=x, out_dtype=torch.float32)
    pad = torch.full((M + 3, C), 7.0, device="cuda")
    pad[:M] = x
    out = ops.swin_mlp(pad[:M], g, b, 1e-5, w1, b1, w2, b2)
    assert relerr(out, ref) < 3e-3, (C, M)
    assert relerr(out, chain) < 3e-3, (C, M)
    assert torch.all(pad[M:] == 7.0)          # rows past M untouched


@pytest.mark.parametrize("M,N,K", [(130, 100, 48), (262, 2304, 768), (1000, 96, 384), (64, 224, 768)])
def test_gemm_simt(cuda, M, N, K):
    from medical_vision_langauge_transformer_b200 import ops
    a, w, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=1 / math.sqrt(K)), rnd(N, seed=3), rnd(M, N, seed=4)
    out = ops.linear(a, w, b, act=1, residual=r)
    ref = F.gelu(a.double() @ w.double().t() + b.double()) + r.double()
    assert relerr(out, ref) < 1e-5


@pytest.mark.parametrize("C", [96, 192, 384, 768, 1536])
@pytest.mark.parametrize("dt_in,dt_out", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                          (torch.bfloat16, torch.bfloat16)])
def test_layernorm(cuda, C, dt_in, dt_out):
    from medical_vision_langauge_transformer_b200 import ops
    x = (rnd(1003, C, seed=C) * 3 + 0.5).to(dt_in)
    g, b = 1 + rnd(C, seed=1, scale=0.1), rnd(C, seed=2, scale=0.1)
    for eps, gelu in ((1e-5, False), (1e-12, False), (1e-5, True)):
        out = ops.layernorm(x, g, b, eps, dt_out, gelu=gelu)
        ref = F.layer_norm(x.float(), (C,), g, b, eps)
        ref = F.gelu(ref) if gelu else ref
        assert relerr(out, ref) < (1e-2 if dt_out == torch.bfloat16 else 2e-6)


def test_layernorm_tiny_variance_eps(cuda):
    """patch-embed outputs of RGC-shaped images have variance ~ eps: eps must be honoured exactly."""
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(64, 96, seed=5, scale=2e-3)
    g, b = torch.ones(96, device="cuda"), torch.zeros(96, device="cuda")
    assert relerr(ops.layernorm(x, g, b, 1e-5, torch.float32), F.layer_norm(x, (96,), g, b, 1e-5)) < 2e-6


def test_patch_embed(cuda):
    from medical_vision_langauge_transformer_b200 import ops
    from oracle import mvlt_oracle as O
    for scale in (1.0, 0.02):
        img = rnd(3, 3, 224, 224, seed=1, scale=scale)
        sd = {"proj.weight": rnd(96, 3, 4, 4, seed=2, scale=0.1), "proj.bias": rnd(96, seed=3, scale=0.02),
              "norm.weight": 1 + rnd(96, seed=4, scale=0.1), "norm.bias": rnd(96, seed=5, scale=0.05)}
        ref = O.patch_embed({k: v.cpu() for k, v in sd.items()}, "", img.cpu())
        for tc in (False, True):   # fp32 CUDA-core kernel; tensor-core kernel with bf16 hi/lo operand splitting
            out = ops.patch_embed_ln(img, sd["proj.weight"], sd["proj.bias"], sd["norm.weight"], sd["norm.bias"],
                                     tensor_cores=tc)
            assert relerr(out.view(3, 3136, 96).cpu(), ref) < 1e-5, (scale, tc)
        g2, b2 = 1 + rnd(96, seed=6, scale=0.1), rnd(96, seed=7, scale=0.05)
        out, a = ops.patch_embed_ln(img, sd["proj.weight"], sd["proj.bias"], sd["norm.weight"], sd["norm.bias"],
                                    tensor_cores=True, next_norm=(g2, b2, 1e-5))
        assert relerr(out.view(3, 3136, 96).cpu(), ref) < 1e-5
        assert relerr(a, F.layer_norm(out, (96,), g2, b2, 1e-5)) < 1e-2          # bf16 second output


@pytest.mark.parametrize("H,C", [(56, 96), (28, 192), (14, 384)])
def test_patch_merge(cuda, H, C):
    from medical_vision_langauge_transformer_b200 import ops
    from oracle import mvlt_oracle as O
    B = 2
    x = rnd(B, H * H, C, seed=H)
    sd = {"norm.weight": 1 + rnd(4 * C, seed=1, scale=0.1), "norm.bias": rnd(4 * C, seed=2, scale=0.05),
          "reduction.weight": rnd(2 * C, 4 * C, seed=3, scale=0.02)}
    ref = O.patch_merging({k: v.cpu() for k, v in sd.items()}, "", x.cpu(), H, H)
    a = ops.patch_merge_ln(x.view(-1, C), sd["norm.weight"], sd["norm.bias"], B, H, H, C, torch.float32)
    out = ops.linear(a, sd["reduction.weight"])
    assert relerr(out.view(B, -1, 2 * C).cpu(), ref) < 1e-5
    a16 = ops.patch_merge_ln(x.view(-1, C), sd["norm.weight"], sd["norm.bias"], B, H, H, C, torch.bfloat16)
    assert relerr(a16, a) < 1e-2


def _window_case(H, C, heads, shift, seed):
    from oracle import mvlt_oracle as O
    B, ws = 2, 7
    qkv = rnd(B * H * H, 3 * C, seed=seed)
    table = rnd(169, heads, seed=seed + 1, scale=0.5)
    # oracle restatement of vfe.py:231-251 given the qkv activations (natural token order)
    x = qkv.cpu().view(B, H, H, 3 * C)
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    xw = O.window_partition(x, ws)                              # [B*nW, 49, 3C]
    q, k, v = xw.view(-1, 49, 3, heads, 32).permute(2, 0, 3, 1, 4)
    attn = (q * 32 ** -0.5) @ k.transpose(-2, -1)
    bias = table.cpu()[O.relative_position_index(ws).reshape(-1)].view(49, 49, heads).permute(2, 0, 1)
    attn = attn + bias[None]
    if shift:
        m = O.shift_attn_mask(H, H, ws, shift)
        nW = m.shape[0]
        attn = (attn.view(B, nW, heads, 49, 49) + m[None, :, None]).view(-1, heads, 49, 49)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, 49, C)
    o = O.window_reverse(o, ws, H, H)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    relb = torch.zeros(heads, 64, 64)
    relb[:, :49, :49] = bias
    return B, qkv, relb.cuda(), o.reshape(B * H * H, C)


@pytest.mark.parametrize("H,C,heads", [(56, 96, 3), (28, 192, 6), (14, 384, 12), (7, 768, 24)])
@pytest.mark.parametrize("shift", [0, 3])
def test_window_attention(cuda, H, C, heads, shift):
    from medical_vision_langauge_transformer_b200 import ops
    if H == 7 and shift:
        pytest.skip("stage 3 never shifts (vfe.py:302-305)")
    B, qkv, relb, ref = _window_case(H, C, heads, shift, seed=H + shift)
    out = ops.window_attention(qkv, relb, B, H, H, C, heads, 7, shift, 32 ** -0.5)
    assert relerr(out.cpu(), ref) < 1e-5
    frag = ops.window_bias_fragments(relb, shift, 32 ** -0.5)
    out16 = ops.window_attention(qkv.bfloat16(), frag, B, H, H, C, heads, 7, shift, 32 ** -0.5)
    assert relerr(out16.cpu(), ref) < 1.5e-2


@pytest.mark.parametrize("H,C,heads", [(56, 96, 3), (28, 192, 6), (14, 384, 12), (7, 768, 24)])
@pytest.mark.parametrize("shift", [0, 3])
def test_window_attention_tc(cuda, H, C, heads, shift):
    """tcgen05 window attention: qkv rows window-major (the order layernorm_winmajor + the qkv GEMM produce), output natural."""
    from medical_vision_langauge_transformer_b200 import ops
    if H == 7 and shift:
        pytest.skip("stage 3 never shifts (vfe.py:302-305)")
    B, qkv, relb, ref = _window_case(H, C, heads, shift, seed=H + shift)
    perm = ops.window_major_index(B, H, H, 7, shift, device="cuda")
    qkv_wm = torch.empty_like(qkv)
    qkv_wm[perm] = qkv
    table = ops.window_bias_table(relb, shift)
    out16 = ops.window_attention_tc(qkv_wm.bfloat16().contiguous(), table, B, H, H, C, heads, 7, shift, 32 ** -0.5)
    assert relerr(out16.cpu(), ref) < 1.5e-2
    # against the same arithmetic on the bf16-rounded operands: only P's bf16 rounding and the fp32 sums differ
    out_ref16 = ops.window_attention(qkv.bfloat16().float(), relb, B, H, H, C, heads, 7, shift, 32 ** -0.5)
    assert relerr(out16, out_ref16) < 6e-3


@pytest.mark.parametrize("B,H,C,heads,shift", [(1, 7, 768, 24, 0), (3, 7, 768, 24, 0), (5, 14, 384, 12, 3), (64, 14, 384, 12, 3),
                                              (64, 14, 384, 12, 0), (16, 56, 96, 3, 3), (33, 28, 192, 6, 3)])
def test_window_attention_tc_batches(cuda, B, H, C, heads, shift):
    """odd window counts (a tile with one window), many tiles per CTA (slot reuse), every stage geometry; checked against
    the fp32 kernel on the bf16-rounded operands."""
    from medical_vision_langauge_transformer_b200 import ops
    qkv = rnd(B * H * H, 3 * C, seed=B + H).bfloat16()
    relb = torch.zeros(heads, 64, 64, device="cuda")
    relb[:, :49, :49] = rnd(heads, 49, 49, seed=7, scale=0.5)
    ref = ops.window_attention(qkv.float(), relb, B, H, H, C, heads, 7, shift, 32 ** -0.5)
    perm = ops.window_major_index(B, H, H, 7, shift, device="cuda")
    qkv_wm = torch.empty_like(qkv)
    qkv_wm[perm] = qkv
    out = ops.window_attention_tc(qkv_wm, ops.window_bias_table(relb, shift), B, H, H, C, heads, 7, shift, 32 ** -0.5)
    assert torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 6e-3
    out2 = ops.window_attention_tc(qkv_wm, ops.window_bias_table(relb, shift), B, H, H, C, heads, 7, shift, 32 ** -0.5)
    assert torch.equal(out, out2), "run-to-run bit reproducibility"


def _tail_case(M, C, seed):
    x = rnd(M, C, seed=seed)
    o = rnd(M, C, seed=seed + 1).bfloat16()
    wp, bp = rnd(C, C, seed=seed + 2, scale=C ** -0.5).bfloat16(), rnd(C, seed=seed + 3, scale=0.1)
    g, b = 1 + rnd(C, seed=seed + 4, scale=0.1), rnd(C, seed=seed + 5, scale=0.1)
    w1, b1 = rnd(4 * C, C, seed=seed + 6, scale=C ** -0.5).bfloat16(), rnd(4 * C, seed=seed + 7, scale=0.1)
    w2, b2 = rnd(C, 4 * C, seed=seed + 8, scale=(4 * C) ** -0.5).bfloat16(), rnd(C, seed=seed + 9, scale=0.1)
    return x, o, wp, bp, g, b, w1, b1, w2, b2


@pytest.mark.parametrize("M,C", [(256, 384), (12544, 384), (300, 384), (37, 384), (1568, 192), (50176, 192), (129, 192), (6272, 96),
                                 (200704, 96), (77, 96), (19000, 96), (40000, 96), (57000, 96)])   # C = 96: 1 .. 11 tiles per persistent CTA pair
@pytest.mark.parametrize("with_proj", [True, False])
def test_swin_block_tail(cuda, M, C, with_proj):
    """proj + residual + LayerNorm + fc1 + GELU + fc2 + residual in one CTA-pair kernel against (1) a torch restatement with
    the kernel's rounding points (bf16 LayerNorm output and hidden activation) and (2) the unfused kernel chain."""
    from medical_vision_langauge_transformer_b200 import ops
    x, o, wp, bp, g, b, w1, b1, w2, b2 = _tail_case(M, C, seed=C + M)
    x1 = x + o.float() @ wp.float().t() + bp if with_proj else x.clone()
    a = F.layer_norm(x1, (C,), g, b, 1e-5).bfloat16().float()
    h = F.gelu(a @ w1.float().t() + b1).bfloat16().float()
    ref = x1 + h @ w2.float().t() + b2
    # unfused chain on the same kernels the model used before
    xc = x.clone()
    if with_proj:
        ops.linear(o, wp, bp, residual=xc, out=xc)
    an = ops.layernorm(xc, g, b, 1e-5, torch.bfloat16)
    hn = ops.linear(an, w1, b1, act=ops.ACT_GELU)
    ops.linear(hn, w2, b2, residual=xc, out=xc)
    xf = x.clone()
    ops.swin_block_tail(xf, o if with_proj else None, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
    assert torch.isfinite(xf).all()
    assert relerr(xf, ref) < 3e-3, (M, C, relerr(xf, ref))
    assert relerr(xf, xc) < 3e-3, (M, C, relerr(xf, xc))
    xf2 = x.clone()
    ops.swin_block_tail(xf2, o if with_proj else None, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
    assert torch.equal(xf, xf2), "run-to-run bit reproducibility"


@pytest.mark.parametrize("M", [300, 57000])
def test_swin_block_tail_c96_both_kernels_agree(cuda, monkeypatch, M):
    """Stage 0 has two block-tail kernels: the persistent three-tiles-in-flight one (default) and the one-tile-per-pair one
    (MVLT_TAIL96=0).  Same arithmetic except for the LayerNorm statistics (one shifted sweep vs two passes): fp32 results agree
    to rounding."""
    from medical_vision_langauge_transformer_b200 import ops
    x, o, wp, bp, g, b, w1, b1, w2, b2 = _tail_case(M, 96, 11)
    xa, xb = x.clone(), x.clone()
    ops.swin_block_tail(xa, o, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
    monkeypatch.setenv("MVLT_TAIL96", "0")
    ops.swin_block_tail(xb, o, wp, bp, g, b, 1e-5, w1, b1, w2, b2)
    assert relerr(xa, xb) < 2e-3, relerr(xa, xb)                           # bf16 roundings of LayerNorm rows that differ in the last fp32 bit
    assert not torch.equal(xa, x)


@pytest.mark.parametrize("B,H,C,shift", [(2, 14, 384, 0), (2, 14, 384, 3), (64, 14, 384, 3), (3, 28, 192, 3), (64, 28, 192, 0), (1, 14, 384, 3),
                                           (1, 56, 96, 0), (2, 56, 96, 3), (64, 56, 96, 3)])
def test_swin_ln_qkv(cuda, B, H, C, shift):
    """norm1 + roll + window_partition + qkv in one CTA-pair kernel against the two-kernel path (window-major LayerNorm, then the
    tcgen05 GEMM): same bf16 A operand, same k order -> the results agree to fp32 summation order; and against torch."""
    from medical_vision_langauge_transformer_b200 import ops
    N = 3 * C
    x = rnd(B * H * H, C, seed=B + H)
    g, b = 1 + rnd(C, seed=4, scale=0.1), rnd(C, seed=5, scale=0.1)
    w, bias = rnd(N, C, seed=6, scale=C ** -0.5).bfloat16(), rnd(N, seed=7, scale=0.1)
    a = ops.layernorm_winmajor(x, g, b, 1e-5, B, H, H, 7, shift)
    two = ops.linear(a, w, bias)
    one = ops.swin_ln_qkv(x, g, b, 1e-5, w, bias, B, H, H, 7, shift)
    assert one.shape == two.shape and one.dtype == torch.bfloat16
    assert relerr(one, two) < 4e-3, relerr(one, two)               # one bf16 ulp where the fp32 sums differ in the last bit
    perm = ops.window_major_index(B, H, H, 7, shift, device="cuda")
    ref = torch.empty(B * H * H, N, device="cuda")
    ref[perm] = F.layer_norm(x, (C,), g, b, 1e-5).bfloat16().float() @ w.float().t() + bias
    assert relerr(one, ref) < 4e-3
    assert torch.equal(one, ops.swin_ln_qkv(x, g, b, 1e-5, w, bias, B, H, H, 7, shift))


@pytest.mark.parametrize("M,K", [(1, 768), (77, 768), (128, 3072), (257, 768), (300, 3072), (8384, 768), (8384, 3072), (20000, 784), (70000, 768)])
def test_linear_residual_layernorm(cuda, M, K):
    """dense + bias + residual + LayerNorm in one cluster-of-four kernel (BertSelfOutput / BertOutput) against torch fp32 on the same
    bf16 operands, against the two-kernel path (tcgen05 GEMM with reduce-add epilogue, then layernorm_rows), in place, ragged M
    (row tiles of 256), a K that is not a multiple of the 64-wide k-block, several tiles per cluster (M = 70000), bit reproducible."""
    from medical_vision_langauge_transformer_b200 import ops
    N = 768
    a = rnd(M, K, seed=1).bfloat16(); res = rnd(M, N, seed=2, scale=2.0) + 0.5
    w, bias = rnd(N, K, seed=3, scale=K ** -0.5).bfloat16(), rnd(N, seed=4, scale=0.1)
    g, b = 1 + rnd(N, seed=5, scale=0.1), rnd(N, seed=6, scale=0.1)
    ref = F.layer_norm(a.float() @ w.float().t() + bias + res, (N,), g, b, 1e-12)
    out, shadow = ops.linear_residual_layernorm(a, w, bias, res, g, b, 1e-12)
    assert out.dtype == torch.float32 and shadow.dtype == torch.bfloat16 and out.shape == shadow.shape == (M, N)
    assert relerr(out, ref) < 1e-5, relerr(out, ref)                       # fp32 accumulation-order noise only
    assert torch.equal(shadow, out.bfloat16())                              # the shadow is the fp32 row rounded once
    two = res.clone()
    ops.linear(a, w, bias, residual=two, out=two)
    two, two_b = ops.layernorm(two, g, b, 1e-12, torch.float32, out=two, bf16_copy=True)
    assert relerr(out, two) < 1e-5
    r2 = res.clone()                                                        # in place on the residual stream
    o2, s2 = ops.linear_residual_layernorm(a, w, bias, r2, g, b, 1e-12, out=r2)
    assert o2.data_ptr() == r2.data_ptr() and torch.equal(o2, out) and torch.equal(s2, shadow)
    o3 = ops.linear_residual_layernorm(a, w, None, res, g, b, 1e-12, bf16_copy=False)   # no bias, no shadow
    assert relerr(o3, F.layer_norm(a.float() @ w.float().t() + res, (N,), g, b, 1e-12)) < 1e-5


def test_linear_residual_layernorm_unsupported_width(cuda):
    from medical_vision_langauge_transformer_b200 import ops
    a, w = rnd(64, 256).bfloat16(), rnd(512, 256).bfloat16()
    with pytest.raises(RuntimeError):
        ops.linear_residual_layernorm(a, w, None, rnd(64, 512), rnd(512), rnd(512), 1e-12)


@pytest.mark.parametrize("B,H,C,shift", [(2, 56, 96, 0), (2, 56, 96, 3), (3, 28, 192, 3), (2, 14, 384, 3), (5, 7, 768, 0)])
def test_layernorm_winmajor(cuda, B, H, C, shift):
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(B * H * H, C, seed=3)
    g, b = 1 + rnd(C, seed=4, scale=0.1), rnd(C, seed=5, scale=0.1)
    nat = ops.layernorm(x, g, b, 1e-5, torch.bfloat16)
    wm = ops.layernorm_winmajor(x, g, b, 1e-5, B, H, H, 7, shift)
    perm = ops.window_major_index(B, H, H, 7, shift, device="cuda")
    assert torch.equal(wm[perm], nat)
    # the index map is the reference's roll + window_partition
    from oracle import mvlt_oracle as O
    tok = torch.arange(B * H * H, dtype=torch.float32).view(B, H, H, 1)
    if shift:
        tok = torch.roll(tok, (-shift, -shift), (1, 2))
    order = O.window_partition(tok, 7).reshape(-1).long()          # window-major position -> natural token
    assert torch.equal(perm.cpu()[order], torch.arange(B * H * H))


def test_patch_embed_window_major_norm(cuda):
    from medical_vision_langauge_transformer_b200 import ops
    B = 3
    img = rnd(B, 3, 224, 224, seed=1, scale=0.02)
    w, b = rnd(96, 3, 4, 4, seed=2, scale=0.1), rnd(96, seed=3, scale=0.1)
    g, be = 1 + rnd(96, seed=4, scale=0.1), rnd(96, seed=5, scale=0.1)
    g2, be2 = 1 + rnd(96, seed=6, scale=0.1), rnd(96, seed=7, scale=0.1)
    out, nat = ops.patch_embed_ln(img, w, b, g, be, 1e-5, tensor_cores=True, next_norm=(g2, be2, 1e-5))
    out_w, wm = ops.patch_embed_ln(img, w, b, g, be, 1e-5, tensor_cores=True, next_norm=(g2, be2, 1e-5), next_norm_window=7)
    assert torch.equal(out, out_w)
    perm = ops.window_major_index(B, 56, 56, 7, 0, device="cuda")
    assert torch.equal(wm[perm], nat)


@pytest.mark.parametrize("B,S", [(3, 131), (64, 131), (1, 131), (7, 74), (64, 74), (5, 81), (2, 96), (2, 64), (3, 128), (3, 144)])
@pytest.mark.parametrize("seq2seq", [False, True])
def test_joint_attention_tc(cuda, B, S, seq2seq):
    """tcgen05 joint attention (tiles of 128 consecutive rows spanning several samples under the lane mask) against the fp32
    kernel on the bf16-rounded operands, with ragged key masks."""
    from medical_vision_langauge_transformer_b200 import ops
    heads, D = 12, 768
    qkv = rnd(B * S, 3 * D, seed=S + B).bfloat16()
    g = torch.Generator().manual_seed(S)
    valid = torch.randint(52, S + 1, (B,), generator=g)
    kmask = torch.where(torch.arange(S)[None] < valid[:, None], 0.0, -10000.0).float().cuda().contiguous()
    ref = ops.joint_attention(qkv.float(), kmask, B, S, heads, seq2seq, 50)
    out = ops.joint_attention(qkv, kmask, B, S, heads, seq2seq, 50, impl="tc")
    assert torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 6e-3
    old = ops.joint_attention(qkv, kmask, B, S, heads, seq2seq, 50, impl="warp")
    assert relerr(out, old) < 6e-3
    assert torch.equal(out, ops.joint_attention(qkv, kmask, B, S, heads, seq2seq, 50, impl="tc")), "run-to-run bit reproducibility"


@pytest.mark.parametrize("L", [80, 23, 30, 1])
@pytest.mark.parametrize("seq2seq", [False, True])
def test_joint_embed_and_attention(cuda, L, seq2seq):
    from medical_vision_langauge_transformer_b200 import ops, synth
    from oracle import mvlt_oracle as O
    B, D, heads = 3, 768, 12
    S = 51 + L
    feat = rnd(B, 49, D, seed=1)
    ids = synth.synth_token_ids(B, L, seed=2, min_len=1).cuda()
    sd = {"MVLBert.word_embeddings.weight": rnd(30523, D, seed=3), "MVLBert.position_embeddings.weight": rnd(512, D, seed=4),
          "MVLBert.token_type_embeddings.weight": rnd(3, D, seed=5)}
    pos = torch.arange(S, device="cuda")
    typepos = sd["MVLBert.token_type_embeddings.weight"][(pos <= 50).long()] + sd["MVLBert.position_embeddings.weight"][pos]
    h, _, kmask = ops.joint_embed(feat, ids, ids > 0, None, sd["MVLBert.word_embeddings.weight"], typepos.contiguous(), 101, 102)
    ref_h = O.joint_embedding({k: v.cpu() for k, v in sd.items()}, ids.cpu(), feat.cpu())
    assert relerr(h.view(B, S, D).cpu(), ref_h) < 1e-6
    ref_mask = O.joint_attention_mask(ids.cpu(), 49, False)[:, 0, 0]
    assert torch.equal(kmask.cpu(), ref_mask)
    # attention on random qkv
    qkv = rnd(B * S, 3 * D, seed=6)
    q, k, v = qkv.cpu().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    mask = O.joint_attention_mask(ids.cpu(), 49, seq2seq)
    ref = (((q @ k.transpose(2, 3)) * 0.125 + mask).softmax(-1) @ v).transpose(1, 2).reshape(B * S, D)
    out = ops.joint_attention(qkv, kmask, B, S, heads, seq2seq, 50)
    assert relerr(out.cpu(), ref) < 1e-5
    out16 = ops.joint_attention(qkv.bfloat16(), kmask, B, S, heads, seq2seq, 50)
    assert relerr(out16.cpu(), ref) < 1.5e-2


@pytest.mark.parametrize("rows,N", [(2560, 30522), (160, 30522), (37, 1000), (300, 256), (5, 33)])
def test_mlm_ce_fused(cuda, rows, N):
    """vocabulary GEMM + online-logsumexp epilogue + finishing pass against F.cross_entropy on the fp32 logits of the same bf16
    operands: ragged vocabulary (30522 = 119 tiles of 256 + 58 columns), ignored rows, deterministic result."""
    from medical_vision_langauge_transformer_b200 import ops
    K = 768
    t = rnd(rows, K, seed=1).bfloat16()
    w = rnd(N, K, seed=2, scale=2 * K ** -0.5).bfloat16()
    b = rnd(N, seed=3, scale=0.5)
    g = torch.Generator().manual_seed(4)
    labels = torch.randint(0, N, (rows,), generator=g)
    labels[torch.rand(rows, generator=g) < 0.8] = -100
    labels[0] = N - 1                                    # a label in the ragged last tile
    labels = labels.cuda()
    logits = t.float() @ w.float().t() + b
    ref_sum = F.cross_entropy(logits, labels, ignore_index=-100, reduction="sum").item()
    n_lab = int((labels != -100).sum())
    acc = ops.mlm_ce_fused(t, w, b, labels, -100)
    assert acc[1].item() == n_lab
    assert abs(acc[0].item() - ref_sum) < 2e-4 * abs(ref_sum), (acc[0].item(), ref_sum)
    acc2 = ops.mlm_ce_fused(t, w, b, labels, -100)
    assert torch.equal(acc, acc2), "fixed-order reduction: bit-identical run to run"
    bad = labels.clone(); bad[0] = N                    # out-of-range label: NaN, not an out-of-bounds read
    assert torch.isnan(ops.mlm_ce_fused(t, w, b, bad, -100)[0])


def test_heads(cuda):
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(37, 768, seed=1)
    w, b = rnd(2, 768, seed=2, scale=0.05), rnd(2, seed=3)
    ref = x @ w.t() + b
    assert relerr(ops.linear_small(x, w, b), ref) < 1e-5
    assert relerr(ops.linear_small(x.bfloat16(), w, b), ref) < 1e-2
    lg = rnd(37, 224, seed=4, scale=3)
    assert relerr(ops.softmax_rows(lg), lg.softmax(-1)) < 1e-5
    labels = torch.randint(0, 224, (37,), device="cuda")
    labels[::3] = -100
    acc = ops.masked_ce(lg, labels, 224)
    ref = F.cross_entropy(lg, labels, ignore_index=-100, reduction="sum")
    assert abs(acc[0].item() - ref.item()) < 1e-3 * abs(ref.item())
    assert acc[1].item() == (labels != -100).sum().item()


@pytest.mark.parametrize("R,C", [(6, 6), (257, 300), (2000, 2000)])
def test_rank_first_positive_device(cuda, R, C):
    """Device-side compute_ranks (run_retrieval.py:220-249) vs the reference's numpy loop: planted diagonal positives plus
    duplicates, lines without any positive, and exact score ties (np.argsort(kind='stable')[::-1] order)."""
    import numpy as np
    from medical_vision_langauge_transformer_b200 import ops
    g = torch.Generator().manual_seed(R * 31 + C)
    scores = torch.rand(R, C, generator=g)
    scores[torch.rand(R, C, generator=g) < 0.2] = 0.5            # ties, also between positives and negatives
    labels = torch.zeros(R, C, dtype=torch.long)
    d = min(R, C)
    labels[torch.arange(d), torch.arange(d)] = 1
    labels[1, min(4, C - 1)] = labels[min(4, R - 1), 1] = 1      # duplicate caption ids (run_retrieval.py:139)
    labels[2, :] = 0                                             # an image without a matching caption
    labels[:, 3] = 0

    def ref(sim, lab):
        out = []
        for s, l in zip(sim, lab):
            hit = np.nonzero(l[np.argsort(s, kind="stable")[::-1]] == 1)[0]
            out.append(int(hit[0]) if hit.size else sim.shape[1])
        return out

    rows, cols = ops.rank_first_positive(scores.cuda(), labels.cuda())
    s, l = scores.numpy(), labels.numpy()
    assert rows.tolist() == ref(s, l)
    assert cols.tolist() == _col_ref(s, l, C)


def _col_ref(s, l, none):
    import numpy as np
    out = []
    for sc, lc in zip(s.T, l.T):
        hit = np.nonzero(lc[np.argsort(sc, kind="stable")[::-1]] == 1)[0]
        out.append(int(hit[0]) if hit.size else none)
    return out


@pytest.mark.parametrize("M,N,K", [(8384, 768, 768), (8384, 768, 3072), (12544, 384, 1536), (12544, 384, 384), (3136, 768, 3072),
                                   (50176, 192, 768), (8200, 768, 320), (5000, 100, 1024)])
def test_gemm_tc_inplace_residual_model_shapes(cuda, M, N, K):
    """In-place residual GEMMs (reduce-add epilogue) at the model's full shapes (BERT attention / FFN outputs, Swin proj / fc2)
    plus ragged M / N / K and a padded row stride: x + a @ w.T + b, columns beyond N untouched."""
    from medical_vision_langauge_transformer_b200 import ops
    a = rnd(M, K, seed=21).bfloat16()
    w = rnd(N, K, seed=22, scale=1 / math.sqrt(K)).bfloat16()
    bias = rnd(N, seed=23)
    ldc = (N + 3) // 4 * 4
    x0 = rnd(M, ldc, seed=24)
    x = x0.clone()
    ops.linear(a, w, bias, residual=x[:, :N], out=x[:, :N])
    ref = x0[:, :N] + a.float() @ w.float().t() + bias
    assert relerr(x[:, :N], ref) < 2e-3
    if ldc > N:
        assert torch.equal(x[:, N:], x0[:, N:])


@pytest.mark.parametrize("C,N,act", [(384, 1152, 0), (384, 1536, 1), (192, 576, 0), (192, 768, 1), (384, 384, 1), (192, 96, 0)])
@pytest.mark.parametrize("M", [12544, 300, 1, 129])
def test_ln_linear_fused(cuda, C, N, act, M):
    """LayerNorm + Linear (+ erf-GELU) in one A-stationary tcgen05 kernel vs (a) torch in fp32 on the bf16-rounded LayerNorm
    output and (b) the unfused LayerNorm kernel + GEMM chain; full tiles, ragged M, a single row, every N chunking
    (1152 = 4.5 x 256, 1536 = 6 x 256, one partial chunk)."""
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(M, C, seed=31) * 1.5 + 0.3
    g, b = 1 + rnd(C, seed=32, scale=0.1), rnd(C, seed=33, scale=0.05)
    w = rnd(N, C, seed=34, scale=1 / math.sqrt(C)).bfloat16()
    bias = rnd(N, seed=35, scale=0.2)
    out = ops.ln_linear(x, g, b, 1e-5, w, bias, act=act)
    a = F.layer_norm(x, (C,), g, b, 1e-5).bfloat16().float()
    ref = a @ w.float().t() + bias
    ref = F.gelu(ref) if act else ref
    assert out.dtype == torch.bfloat16 and out.shape == (M, N)
    assert relerr(out, ref) < 1e-2
    chain = ops.linear(ops.layernorm(x, g, b, 1e-5, torch.bfloat16), w, bias, act=act)
    assert relerr(out, chain) < 1e-2
