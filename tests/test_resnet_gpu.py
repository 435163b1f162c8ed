"""ResNet backbone (BASELINE.json configs[4]) on the GPU: the convolution kernels against torch's own
conv2d / max_pool2d on the same bf16-rounded operands, and the whole ResNet-101 / ResNet-50 + BERT retrieval forward against
golden outputs of the REAL reference (tests/golden/retrieval_resnet*.pt, oracle/make_golden.py).  Tolerances as in
test_e2e_gpu.py: fp32 mode 1e-4 relative, bf16 mode 1e-2 on logits."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def nhwc(x):  # [B,C,H,W] -> [B*H*W, C]
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous()


def tap_major(w):  # [N,C,R,S] -> [N, R*S*C]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


# (B, H, C, N, k, stride, pad): every convolution geometry of the Bottleneck trunk + ragged batch sizes (M % 256 != 0,
# tiles that straddle image boundaries, a single partial tile)
CONV_CASES = [
    (2, 56, 64, 64, 3, 1, 1), (3, 28, 128, 128, 3, 1, 1), (5, 14, 256, 256, 3, 1, 1), (3, 7, 512, 512, 3, 1, 1),
    (2, 56, 128, 128, 3, 2, 1), (3, 28, 256, 256, 3, 2, 1), (5, 14, 512, 512, 3, 2, 1),
    (2, 56, 256, 512, 1, 2, 0), (3, 28, 512, 1024, 1, 2, 0), (5, 14, 1024, 2048, 1, 2, 0),
    (1, 7, 64, 96, 3, 1, 1), (1, 9, 64, 32, 3, 2, 1), (7, 10, 128, 200, 3, 1, 1),
]


@pytest.mark.parametrize("B,H,C,N,k,stride,pad", CONV_CASES)
def test_conv2d_implicit_gemm_bf16(cuda, B, H, C, N, k, stride, pad):
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(B, C, H, H, seed=1).bfloat16()
    w = rnd(N, C, k, k, seed=2, scale=1 / math.sqrt(C * k * k)).bfloat16()
    bias = rnd(N, seed=3, scale=0.1)
    ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=pad)
    out = ops.conv2d_nhwc(nhwc(x), tap_major(w), bias, B, H, H, k, k, stride, pad)
    assert out.dtype == torch.bfloat16 and out.shape == (B * ref.shape[2] * ref.shape[3], N)
    assert relerr(out, nhwc(ref)) < 1e-2
    # + identity + ReLU epilogue (Bottleneck tail) and ReLU alone
    idn = rnd(*ref.shape, seed=4).bfloat16()
    out = ops.conv2d_nhwc(nhwc(x), tap_major(w), bias, B, H, H, k, k, stride, pad, act=ops.ACT_RELU, residual=nhwc(idn))
    assert relerr(out, nhwc(F.relu(ref + idn.float()))) < 1e-2
    out = ops.conv2d_nhwc(nhwc(x), tap_major(w), bias, B, H, H, k, k, stride, pad, act=ops.ACT_RELU)
    assert relerr(out, nhwc(F.relu(ref))) < 1e-2


@pytest.mark.parametrize("B,H,C,N,k,stride,pad", [(2, 14, 64, 96, 3, 1, 1), (3, 28, 32, 40, 3, 2, 1), (2, 56, 64, 128, 1, 2, 0)])
def test_conv2d_fp32_parity_mode(cuda, B, H, C, N, k, stride, pad):
    from medical_vision_langauge_transformer_b200 import ops
    x = rnd(B, C, H, H, seed=1)
    w = rnd(N, C, k, k, seed=2, scale=1 / math.sqrt(C * k * k))
    bias = rnd(N, seed=3, scale=0.1)
    # fp64 reference: cuDNN's fp32 convolution runs on TF32 tensor cores by default (1e-3 relative), not a 1e-4 yardstick
    ref = F.conv2d(x.double(), w.double(), bias.double(), stride=stride, padding=pad).float()
    idn = rnd(*ref.shape, seed=4)
    out = ops.conv2d_nhwc(nhwc(x), tap_major(w), bias, B, H, H, k, k, stride, pad, act=ops.ACT_RELU, residual=nhwc(idn))
    assert relerr(out, nhwc(F.relu(ref + idn))) < 1e-4


@pytest.mark.parametrize("act,res", [(3, False), (3, True), (4, True), (0, True), (4, False)])
def test_gemm_tc_relu_epilogues_bf16_residual(cuda, act, res):
    """1x1 convolutions of the trunk: relu(A.W^T + b (+ identity)) [then GELU for the last block], bf16 in / bf16 out."""
    from medical_vision_langauge_transformer_b200 import ops
    M, N, K = 1000, 1024, 256
    a = rnd(M, K, seed=5).bfloat16()
    w = rnd(N, K, seed=6, scale=1 / math.sqrt(K)).bfloat16()
    bias = rnd(N, seed=7)
    r = rnd(M, N, seed=8).bfloat16() if res else None
    out = ops.linear(a, w, bias, act=act, residual=r, out_dtype=torch.bfloat16)
    ref = a.float() @ w.float().t() + bias
    if res:
        ref = ref + r.float()
    if act >= 3:
        ref = F.relu(ref)
    if act == 4:
        ref = F.gelu(ref)
    assert relerr(out, ref) < 1e-2
    if res:  # in place on the identity buffer
        r2 = r.clone()
        ops.linear(a, w, bias, act=act, residual=r2, out=r2)
        assert torch.equal(r2, out)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem_im2col_and_maxpool(cuda, dtype):
    from medical_vision_langauge_transformer_b200 import ops
    B = 2
    img = rnd(B, 3, 224, 224, seed=9)
    p = ops.stem_im2col(img, 7, 7, 2, 3, 160, dtype)
    ref = F.unfold(img, 7, padding=3, stride=2).transpose(1, 2).reshape(B * 112 * 112, 147)   # k = (c*7 + ky)*7 + kx
    assert p.shape == (B * 112 * 112, 160) and p.dtype == dtype
    assert torch.equal(p[:, :147].float(), ref.to(dtype).float()) and p[:, 147:].abs().max().item() == 0
    x = rnd(B, 64, 112, 112, seed=10).to(dtype)
    out = ops.maxpool_nhwc(nhwc(x), B, 112, 112, 3, 2, 1)
    assert torch.equal(out.float(), nhwc(F.max_pool2d(x.float(), 3, 2, 1)))


@pytest.mark.parametrize("B,H,W", [(2, 224, 224), (1, 64, 96), (3, 46, 30)])
def test_resnet_stem_fused(cuda, B, H, W):
    """conv1 7x7/2 + folded bn1 + relu + maxpool 3x3/2 in one kernel vs torch on the same bf16-rounded operands; ragged
    sizes exercise partial pooled tiles and the halo rows / columns outside the image."""
    from medical_vision_langauge_transformer_b200 import ops
    img = rnd(B, 3, H, W, seed=12)
    w = rnd(64, 3, 7, 7, seed=13, scale=1 / math.sqrt(147))
    bias = rnd(64, seed=14, scale=0.2)
    wp = torch.zeros(64, 160, device="cuda")
    wp[:, :147] = w.reshape(64, 147)
    out, Hp, Wp = ops.resnet_stem(img, wp.bfloat16().contiguous(), bias)
    conv = F.conv2d(img.bfloat16().float(), w.bfloat16().float(), bias, stride=2, padding=3)
    ref = F.max_pool2d(F.relu(conv).bfloat16().float(), 3, 2, 1)
    assert (Hp, Wp) == tuple(ref.shape[2:]) and out.shape == (B * Hp * Wp, 64)
    assert relerr(out, nhwc(ref)) < 1e-2


def test_im2col_explicit(cuda):
    """The parity-mode patch matrix is tap-major: k = (ky*S + kx)*C + c."""
    from medical_vision_langauge_transformer_b200 import _lib, ops
    lib = _lib.ensure_init()
    B, C, H, k, stride, pad = 2, 32, 13, 3, 2, 1
    for dtype, code in ((torch.float32, 0), (torch.bfloat16, 1)):
        x = rnd(B, C, H, H, seed=11).to(dtype)
        Ho = (H + 2 * pad - k) // stride + 1
        out = torch.empty(B * Ho * Ho, k * k * C, device="cuda", dtype=dtype)
        rc = lib.mvlt_im2col_nhwc(nhwc(x).data_ptr(), code, out.data_ptr(), k * k * C, B, H, H, C, k, k, stride, pad,
                                  torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        u = F.unfold(x.float(), k, padding=pad, stride=stride)                  # [B, C*k*k, L], k index = c*k*k + tap
        ref = u.view(B, C, k * k, -1).permute(0, 3, 2, 1).reshape(B * Ho * Ho, k * k * C)
        assert torch.equal(out.float(), ref)


def build(conv, precision, L=80):
    from medical_vision_langauge_transformer_b200 import synth
    from medical_vision_langauge_transformer_b200.modules import config as C, model as M
    model = M.MVLBertForRetrieval(C.offline_config("retrieval", conv=conv, max_length=L)).eval()
    sd = {k: v.clone() for k, v in synth.load_synth(model, 0, "stress").items()}
    return model.cuda().set_precision(precision), sd


def probe_check(taps, golden_taps, tol, label):
    from oracle.make_golden import probe_indices
    worst = {}
    for name, g in golden_taps.items():
        if name not in taps:
            continue
        t = taps[name].detach().float().cpu().contiguous()
        assert tuple(t.shape) == tuple(g["shape"]), (name, t.shape, g["shape"])
        vals = t.flatten()[probe_indices(t.numel(), name)]
        worst[name] = ((vals - g["values"]).abs().max() / max(g["absmax"], 1e-12)).item()
    bad = {k: v for k, v in worst.items() if v > tol}
    assert not bad, f"{label}: probes out of tolerance {bad} (all: {worst})"
    return worst


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("conv", ["resnet101", "resnet50"])
def test_resnet_retrieval_vs_reference_golden(cuda, conv, precision):
    from medical_vision_langauge_transformer_b200 import synth
    g = torch.load(os.path.join(GOLDEN, f"retrieval_{conv}.pt"))
    assert g["conv"] == conv
    model, _ = build(conv, precision, g["L"])
    x = synth.synth_images(g["B"], g["data_seed"], g["img_scale"]).cuda()
    ids = synth.synth_token_ids(g["B"], g["L"], g["data_seed"]).cuda()
    taps = {}
    model.conv.conv[0].taps = model.MVLBert.taps = taps
    with torch.no_grad():
        prob = model(x, ids)
        logits = model(x, ids, image_text_label=torch.zeros(g["B"], dtype=torch.long, device="cuda"))
    assert {"stem", "layer1", "layer2", "layer3", "image_feature", "bert11"} <= set(taps)
    worst = probe_check(taps, g["taps"], 1e-4 if precision == "fp32" else 3e-2, f"{conv}/{precision}")
    print(conv, precision, {k: f"{v:.1e}" for k, v in worst.items()}, "logits", relerr(logits, g["logits"]))
    if precision == "fp32":
        assert relerr(logits, g["logits"]) < 1e-4 and relerr(prob, g["prob"]) < 1e-4
    else:
        assert torch.allclose(logits.cpu(), g["logits"], rtol=1e-2, atol=1e-2), (logits, g["logits"])
        assert relerr(prob, g["prob"]) < 1e-2
    assert prob.dtype == torch.float32 and prob.shape == (g["B"], 2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("conv", ["linear", "vit"])
def test_linear_and_vit_retrieval_vs_reference_golden(cuda, conv, precision):
    """config.conv = 'linear' (vfe.py:47-60) and 'vit' (vfe.py:66-107): 196 image tokens, S = 278 through the 288-key joint
    attention; the ViT encoder blocks run on the BERT kernels (197-token attention, no mask)."""
    from medical_vision_langauge_transformer_b200 import synth
    g = torch.load(os.path.join(GOLDEN, f"retrieval_{conv}.pt"))
    model, _ = build(conv, precision, g["L"])
    x = synth.synth_images(g["B"], g["data_seed"], g["img_scale"]).cuda()
    ids = synth.synth_token_ids(g["B"], g["L"], g["data_seed"]).cuda()
    taps = {}
    model.conv.conv[0].taps = model.MVLBert.taps = taps
    with torch.no_grad():
        prob = model(x, ids)
        logits = model(x, ids, image_text_label=torch.zeros(g["B"], dtype=torch.long, device="cuda"))
    assert taps["image_feature"].shape == (g["B"], 196, 768)
    worst = probe_check(taps, g["taps"], 1e-4 if precision == "fp32" else 3e-2, f"{conv}/{precision}")
    print(conv, precision, {k: f"{v:.1e}" for k, v in worst.items()}, "logits", relerr(logits, g["logits"]))
    if precision == "fp32":
        assert relerr(logits, g["logits"]) < 1e-4 and relerr(prob, g["prob"]) < 1e-4
    else:
        assert torch.allclose(logits.cpu(), g["logits"], rtol=1e-2, atol=1e-2), (logits, g["logits"])
        assert relerr(prob, g["prob"]) < 1e-2


def test_resnet_backbone_module_surface(cuda):
    """`resnet101_without_fc()(x)` keeps the reference's output contract (vfe.py:14-24): NCHW [B, 2048, 7, 7]; checked
    against the oracle trunk on another seed, and train mode is rejected (no batch-statistics kernel)."""
    from medical_vision_langauge_transformer_b200 import synth
    from medical_vision_langauge_transformer_b200.modules.visual_feature_extractor import resnet50_without_poolfc
    from oracle import mvlt_oracle as O
    net = resnet50_without_poolfc(precision="fp32").eval()
    sd = {("conv.conv.0." + k): v for k, v in net.state_dict().items()}
    new = synth.synth_state_dict({k: v.shape for k, v in sd.items()}, 5, "stress")
    net.load_state_dict({k[len("conv.conv.0."):]: v for k, v in new.items()}, strict=False)
    x = synth.synth_images(2, 7, 1.0)
    with torch.no_grad():
        ref = O.resnet_forward(new, x, "conv.conv.0.", O.RESNET_LAYERS["resnet50"])
        out = net.cuda()(x.cuda())
    assert out.shape == (2, 2048, 7, 7) and relerr(out, ref) < 1e-4
    with pytest.raises(NotImplementedError, match="train-mode BatchNorm"):
        net.train()(x.cuda())
