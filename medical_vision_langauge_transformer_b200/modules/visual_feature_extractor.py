"""Swin visual feature extractor with the module tree / state_dict keys of the reference's
modules/visual_feature_extractor.py:125-702, executed by the sm_100a kernels in libmvlt_b200.so.

The sub-modules below (`PatchEmbed`, `WindowAttention`, `Mlp`, `SwinTransformerBlock`, `PatchMerging`, `BasicLayer`)
only HOLD parameters under the reference's names; the arithmetic of the whole trunk is orchestrated by
`SwinTransformer.forward` so that tokens stay in natural [B, H*W, C] order end to end:

  patch-embed+LN kernel -> per block { LN1 -> qkv GEMM -> window attention (roll/partition/reverse folded into the
  row index map) -> proj GEMM (+bias +residual, in place) -> LN2 -> fc1 GEMM (+bias, erf-GELU) -> fc2 GEMM (+bias
  +residual, in place) } -> per stage { 2x2 gather + LN(4C) -> reduction GEMM } -> final LN (optionally + GELU).

The residual stream is fp32 in both precisions; GEMM operands are bf16 (tcgen05) or fp32 (parity mode).
"""
from __future__ import annotations

import collections
import os

import torch
import torch.nn as nn

from .. import ops

_WS = 7


def default_precision() -> str:
    p = os.environ.get("MVLT_PRECISION", "bf16").lower()
    if p not in ("bf16", "fp32"):
        raise ValueError(f"MVLT_PRECISION must be bf16 or fp32, got {p}")
    return p


def act_dtype(precision: str) -> torch.dtype:
    return torch.bfloat16 if precision == "bf16" else torch.float32


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class Mlp(nn.Module):
    """vfe.py:125-141 parameter holder: fc1 C->hidden, erf-GELU, fc2 hidden->C."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)


class WindowAttention(nn.Module):
    """vfe.py:176-254 parameter holder (bias table [(2w-1)^2, heads], index buffer [w*w, w*w], qkv, proj)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        self.register_buffer("relative_position_index", self.build_relative_position_index())
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        _trunc_normal_(self.relative_position_bias_table)

    def build_relative_position_index(self) -> torch.Tensor:
        """vfe.py:203-214 (input independent; also used to restore the buffer when a checkpoint lacks it)."""
        wh, ww = self.window_size
        ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
        coords = torch.stack([ys.flatten(), xs.flatten()])
        d = coords[:, :, None] - coords[:, None, :]
        return (d[0] + wh - 1) * (2 * ww - 1) + (d[1] + ww - 1)

    def gathered_bias(self) -> torch.Tensor:
        """[heads, 64, 64] fp32, zero padded: table[index] permuted as vfe.py:236-238 (input independent)."""
        n = self.window_size[0] * self.window_size[1]
        b = self.relative_position_bias_table.detach().float()[self.relative_position_index.view(-1)]
        b = b.view(n, n, self.num_heads).permute(2, 0, 1)
        out = torch.zeros(self.num_heads, 64, 64, device=b.device, dtype=torch.float32)
        out[:, :n, :n] = b
        return out.contiguous()


class SwinTransformerBlock(nn.Module):
    """vfe.py:273-387 parameter holder; `attn_mask` buffer kept for state_dict parity (the kernel recomputes the
    region test arithmetically)."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:
            self.shift_size, self.window_size = 0, min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, (self.window_size, self.window_size), num_heads, qkv_bias, qk_scale, attn_drop, drop)
        self.drop_path_rate = drop_path
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.register_buffer("attn_mask", self.build_attn_mask())

    def build_attn_mask(self):
        """vfe.py:318-348: -100 between tokens of different regions of the rolled image (None for unshifted blocks)."""
        if self.shift_size == 0:
            return None
        H, W = self.input_resolution
        ws, sh = self.window_size, self.shift_size
        region = torch.zeros(H, W)
        bounds = ((0, H - ws), (H - ws, H - sh), (H - sh, H))
        for a, (h0, h1) in enumerate(bounds):
            for b, (w0, w1) in enumerate(((0, W - ws), (W - ws, W - sh), (W - sh, W))):
                region[h0:h1, w0:w1] = 3 * a + b
        win = region.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
        diff = win[:, None, :] - win[:, :, None]
        return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class PatchMerging(nn.Module):
    """vfe.py:408-445 parameter holder."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim = tuple(input_resolution), dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)


class BasicLayer(nn.Module):
    """vfe.py:457-513: `depth` blocks alternating shift 0 / window//2, then optional PatchMerging."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth, self.use_checkpoint = dim, tuple(input_resolution), depth, use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, input_resolution, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2,
                                 mlp_ratio, qkv_bias, qk_scale, drop, attn_drop,
                                 drop_path[i] if isinstance(drop_path, (list, tuple)) else drop_path, norm_layer=norm_layer)
            for i in range(depth)])
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None


class PatchEmbed(nn.Module):
    """vfe.py:527-565 parameter holder: Conv2d(in_chans, embed_dim, k=s=patch) + optional LayerNorm."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.patches_resolution = [img_size // patch_size, img_size // patch_size]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None


class SwinTransformer(nn.Module):
    """Same constructor and `forward(x) -> [B, 49, C_last]` as vfe.py:575-693."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, precision=None, **kwargs):
        super().__init__()
        if ape:
            raise NotImplementedError("absolute position embedding (ape=True) is not used by MVLT (model.py:219)")
        if qk_scale is not None:
            raise NotImplementedError("qk_scale override is not used by MVLT (swin yaml leaves it None)")
        depths, num_heads = list(depths), list(num_heads)
        self.num_classes, self.num_layers, self.embed_dim = num_classes, len(depths), embed_dim
        self.ape, self.patch_norm, self.mlp_ratio = ape, patch_norm, mlp_ratio
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.precision = precision or default_precision()
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, norm_layer if patch_norm else None)
        self.patches_resolution = self.patch_embed.patches_resolution
        self.pos_drop = nn.Dropout(p=drop_rate)
        n_blocks = sum(depths)   # torch.linspace(0, rate, n) of vfe.py:633 in plain Python (HF builds models under a meta device)
        dpr = [drop_path_rate * i / max(n_blocks - 1, 1) for i in range(n_blocks)]
        res = self.patches_resolution
        self.layers = nn.ModuleList([
            BasicLayer(int(embed_dim * 2 ** i), (res[0] // 2 ** i, res[1] // 2 ** i), depths[i], num_heads[i], window_size,
                       mlp_ratio, qkv_bias, qk_scale, drop_rate, attn_drop_rate, dpr[sum(depths[:i]):sum(depths[:i + 1])],
                       norm_layer, PatchMerging if i < self.num_layers - 1 else None, use_checkpoint)
            for i in range(self.num_layers)])
        self.norm = norm_layer(self.num_features)
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)
        self._packed = None
        self._packed_key = None
        self.taps = None   # tests set this to a dict to record block-boundary activations (clones)

    @staticmethod
    def _init_weights(m):  # vfe.py:659-666
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)

    # ---------------------------------------------------------------- packed weights
    def _fingerprint(self):
        return (self.precision, self.patch_embed.proj.weight.device,
                sum(p._version for p in self.parameters()), id(self.patch_embed.proj.weight))

    def packed(self):
        """GEMM operands in the layout the kernels want (bf16 copies in bf16 mode) + gathered rel-pos bias; rebuilt
        whenever a parameter is replaced, moved or modified in place."""
        key = self._fingerprint()
        if self._packed is None or self._packed_key != key:
            wd = act_dtype(self.precision)
            f32 = lambda t: t.detach().float().contiguous()
            wcast = lambda t: t.detach().to(wd).contiguous()
            pk = {"pe_w": f32(self.patch_embed.proj.weight), "pe_b": f32(self.patch_embed.proj.bias), "blocks": [], "merge": []}
            for layer in self.layers:
                for blk in layer.blocks:
                    pk["blocks"].append(dict(
                        n1w=f32(blk.norm1.weight), n1b=f32(blk.norm1.bias), n2w=f32(blk.norm2.weight), n2b=f32(blk.norm2.bias),
                        qkv_w=wcast(blk.attn.qkv.weight), qkv_b=f32(blk.attn.qkv.bias),
                        proj_w=wcast(blk.attn.proj.weight), proj_b=f32(blk.attn.proj.bias),
                        fc1_w=wcast(blk.mlp.fc1.weight), fc1_b=f32(blk.mlp.fc1.bias),
                        fc2_w=wcast(blk.mlp.fc2.weight), fc2_b=f32(blk.mlp.fc2.bias),
                        relbias=(ops.window_bias_fragments(blk.attn.gathered_bias(), blk.shift_size, blk.attn.scale,
                                                           blk.window_size)
                                 if self.precision == "bf16" else blk.attn.gathered_bias()),
                        bias_tc=(ops.window_bias_table(blk.attn.gathered_bias(), blk.shift_size, blk.window_size)
                                 if self.precision == "bf16" else None)))
                if layer.downsample is not None:
                    pk["merge"].append(dict(nw=f32(layer.downsample.norm.weight), nb=f32(layer.downsample.norm.bias),
                                            red_w=wcast(layer.downsample.reduction.weight)))
            self._packed, self._packed_key = pk, key
        return self._packed

    # ---------------------------------------------------------------- forward
    def forward_features(self, x, final_gelu: bool = False, out_dtype=None):
        pe = self.patch_embed
        if self.training and (any(getattr(b, "drop_path_rate", 0.0) > 0 for l in self.layers for b in l.blocks) or self.pos_drop.p > 0):
            raise NotImplementedError("train mode (DropPath / Dropout, vfe.py:313,:633) is outside the accelerated forward path; "
                                      "call model.eval() — this package implements the eval-mode forward only")
        B, C_in, H_img, W_img = x.shape
        assert H_img == pe.img_size[0] and W_img == pe.img_size[1], \
            f"Input image size ({H_img}*{W_img}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]})."
        if not x.is_cuda:
            raise RuntimeError("mvlt_b200 SwinTransformer runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        pk = self.packed()
        adt = act_dtype(self.precision)
        x = x.contiguous().float()
        a_first = None
        # bf16 mode: window attention on tcgen05 (MVLT_ATTN=tc, default) reads qkv rows WINDOW-MAJOR: norm1 writes its bf16
        # output in that order (roll + window_partition cost nothing), the qkv GEMM keeps it, the attention kernel scatters
        # its output back to natural order.  MVLT_ATTN=warp: the mma.sync kernel of round 1 on natural-order rows.
        attn_tc = self.precision == "bf16" and ops.attention_impl() == "tc"
        if self.precision == "bf16":   # tensor-core stem; norm1 of block 0 comes out of the same kernel
            b0 = self.layers[0].blocks[0]
            X, a_first = ops.patch_embed_ln(x, pk["pe_w"], pk["pe_b"], pe.norm.weight, pe.norm.bias, pe.norm.eps,
                                            tensor_cores=True, next_norm=(pk["blocks"][0]["n1w"], pk["blocks"][0]["n1b"], b0.norm1.eps),
                                            next_norm_window=b0.window_size if (attn_tc and b0.window_size == 7) else 0)
        else:
            X = ops.patch_embed_ln(x, pk["pe_w"], pk["pe_b"], pe.norm.weight, pe.norm.bias, pe.norm.eps)
        # LN2 + fc1 + GELU + fc2 + residual as ONE tcgen05 kernel (csrc/swin_mlp.cu) for the stage widths listed in
        # MVLT_FUSED_MLP (bf16 mode).  Default: stage 0 only (C = 96, HBM-bound: 96 us vs 136 us for the unfused chain);
        # at C = 192 / 384 the kernel is shared-memory-bandwidth bound at its N = 64 MMAs and the unfused chain is faster
        # (profiles/r01_fused_mlp_vs_chain.log), "96,192,384" enables it everywhere, "" nowhere.
        fused_widths = tuple(int(v) for v in os.environ.get("MVLT_FUSED_MLP", "96").split(",") if v.strip()) \
            if self.precision == "bf16" else ()
        # LayerNorm + Linear as ONE A-stationary tcgen05 kernel (norm1 -> qkv, norm2 -> fc1 + GELU) for the widths listed in
        # MVLT_FUSED_LN_LINEAR (bf16 mode).  Default: none — at C = 384 it removes 36 LayerNorm launches but runs 29.9 us per
        # call against 29-33 us for LayerNorm + GEMM (98 row tiles on 148 SMs, N = 256 chunks fed from shared memory at ~70 %
        # of the tensor rate): 13.87 k vs 14.00 k pairs/s for the step (profiles/r01_ln_linear_fused_experiment.log).
        ln_lin_widths = tuple(int(v) for v in os.environ.get("MVLT_FUSED_LN_LINEAR", "").split(",") if v.strip()) \
            if self.precision == "bf16" else ()
        # proj + residual + LN2 + fc1 + GELU + fc2 + residual as ONE tcgen05 kernel on CTA pairs (csrc/swin_tail.cu; the new residual
        # rows stay in tensor memory between the two halves) for the widths in MVLT_BLOCK_TAIL (bf16 mode; default 96,192,384 — stage 0 runs the persistent
        # two-tiles-in-flight variant csrc/swin_tail96.cu: 72 us against 140 us for proj GEMM + swin_mlp_fused).
        tail_widths = tuple(int(v) for v in os.environ.get("MVLT_BLOCK_TAIL", "96,192,384").split(",") if v.strip()) \
            if self.precision == "bf16" else ()
        # norm1 + roll + window_partition + qkv as ONE tcgen05 kernel on CTA pairs (csrc/ln_qkv.cu) for the widths in MVLT_LN_QKV
        lnqkv_widths = tuple(int(v) for v in os.environ.get("MVLT_LN_QKV", "192,384").split(",") if v.strip()) \
            if self.precision == "bf16" else ()
        taps = self.taps
        if taps is not None:
            taps["patch_embed"] = X.clone().view(B, -1, X.shape[-1])
        bi = 0
        for s, layer in enumerate(self.layers):
            H, W = layer.input_resolution
            C = layer.dim
            assert X.shape == (B * H * W, C), "input feature has wrong size"
            for i, blk in enumerate(layer.blocks):
                w = pk["blocks"][bi]
                bi += 1
                ln_lin = C in ln_lin_widths and C in ops.FUSED_LN_LINEAR_WIDTHS
                win_tc = attn_tc and blk.window_size == 7 and H % 7 == 0 and W % 7 == 0
                if a_first is not None:
                    a, a_first = a_first, None
                    qkv = ops.linear(a, w["qkv_w"], w["qkv_b"])
                elif win_tc and C in lnqkv_widths and C in ops.LN_QKV_WIDTHS and ops.use_ln_qkv(B * H * W, C):
                    qkv = ops.swin_ln_qkv(X, w["n1w"], w["n1b"], blk.norm1.eps, w["qkv_w"], w["qkv_b"], B, H, W, blk.window_size,
                                          blk.shift_size)
                elif win_tc:
                    a = ops.layernorm_winmajor(X, w["n1w"], w["n1b"], blk.norm1.eps, B, H, W, blk.window_size, blk.shift_size)
                    qkv = ops.linear(a, w["qkv_w"], w["qkv_b"])
                elif ln_lin:
                    qkv = ops.ln_linear(X, w["n1w"], w["n1b"], blk.norm1.eps, w["qkv_w"], w["qkv_b"])
                else:
                    a = ops.layernorm(X, w["n1w"], w["n1b"], blk.norm1.eps, adt)
                    qkv = ops.linear(a, w["qkv_w"], w["qkv_b"])
                if win_tc:
                    o = ops.window_attention_tc(qkv, w["bias_tc"], B, H, W, C, blk.num_heads, blk.window_size, blk.shift_size,
                                                blk.attn.scale)
                else:
                    o = ops.window_attention(qkv, w["relbias"], B, H, W, C, blk.num_heads, blk.window_size, blk.shift_size,
                                             blk.attn.scale)
                if C in tail_widths and C in ops.BLOCK_TAIL_WIDTHS and w["fc1_w"].shape[0] == 4 * C and ops.use_block_tail(B * H * W, C):
                    ops.swin_block_tail(X, o, w["proj_w"], w["proj_b"], w["n2w"], w["n2b"], blk.norm2.eps, w["fc1_w"], w["fc1_b"],
                                        w["fc2_w"], w["fc2_b"])
                    if taps is not None and i < 2:
                        taps[f"s{s}b{i}"] = X.clone().view(B, H * W, C)
                    continue
                ops.linear(o, w["proj_w"], w["proj_b"], residual=X, out=X)
                if C in fused_widths and C in ops.FUSED_MLP_WIDTHS and w["fc1_w"].shape[0] == 4 * C:
                    ops.swin_mlp(X, w["n2w"], w["n2b"], blk.norm2.eps, w["fc1_w"], w["fc1_b"], w["fc2_w"], w["fc2_b"])
                elif ln_lin:
                    h = ops.ln_linear(X, w["n2w"], w["n2b"], blk.norm2.eps, w["fc1_w"], w["fc1_b"], act=ops.ACT_GELU)
                    ops.linear(h, w["fc2_w"], w["fc2_b"], residual=X, out=X)
                else:
                    a = ops.layernorm(X, w["n2w"], w["n2b"], blk.norm2.eps, adt)
                    h = ops.linear(a, w["fc1_w"], w["fc1_b"], act=ops.ACT_GELU)
                    ops.linear(h, w["fc2_w"], w["fc2_b"], residual=X, out=X)
                if taps is not None and i < 2:
                    taps[f"s{s}b{i}"] = X.clone().view(B, H * W, C)
            if layer.downsample is not None:
                m = pk["merge"][s]
                assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
                a = ops.patch_merge_ln(X, m["nw"], m["nb"], B, H, W, C, adt, layer.downsample.norm.eps)
                X = ops.linear(a, m["red_w"], out_dtype=torch.float32)
            if taps is not None:
                taps[f"stage{s}"] = X.clone().view(B, -1, X.shape[-1])
        out = ops.layernorm(X, self.norm.weight, self.norm.bias, self.norm.eps, out_dtype or torch.float32, gelu=final_gelu)
        return out.view(B, -1, self.num_features)

    def forward(self, x):
        return self.forward_features(x)


# ====================================================================================================================
# ResNet backbones (vfe.py:7-44): torchvision ResNet-101 / ResNet-50 without pooling and fc, Bottleneck blocks.
# ====================================================================================================================
class Bottleneck(nn.Module):
    """torchvision resnet.py Bottleneck parameter holder (v1.5: the stride sits on conv2): conv1 1x1 -> bn1 -> relu ->
    conv2 3x3/stride -> bn2 -> relu -> conv3 1x1 -> bn3 -> (+ identity | downsample(x)) -> relu."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


def _fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm2d folded into the convolution: y = conv(x; w * s) + (beta - mean * s), s = gamma / sqrt(var + eps).
    -> (weight [N, R*S*C] tap-major fp32, bias [N] fp32)."""
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv.weight.detach().float() * s[:, None, None, None]
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * s
    if conv.bias is not None:
        b = b + conv.bias.detach().float() * s
    return w, b.contiguous()


class ResNetWithoutFC(nn.Module):
    """Module tree / state_dict keys of torchvision.models.ResNet(Bottleneck, layers, num_classes=1000) as subclassed by
    vfe.py:7-24 (`resnet101_without_fc`) and :27-44 (`resnet50_without_poolfc`): `forward(x)` stops after layer4 and returns
    [B, 2048, 7, 7].  `avgpool` / `fc` are constructed (their keys are in reference checkpoints) and never applied.
    Initialisation as torchvision: kaiming-normal (fan_out, relu) convolutions, unit BatchNorm.  Eval-mode BatchNorm only
    (running statistics folded into the GEMM weights): this is the forward path; `train()` statistics are out of scope.

    Arithmetic: activations are NHWC matrices [B*H*W, C] (bf16 in bf16 mode); the stem (conv1 + bn1 + relu + maxpool) is
    one mma.sync kernel, 1x1 convolutions are tcgen05 GEMMs on the matrix as is, 3x3 and strided convolutions are the same GEMM kernel with its
    A operand fetched by im2col-mode TMA, BatchNorm + ReLU + the identity add live in the GEMM epilogue."""

    def __init__(self, layers, pretrained=False, progress=False, precision=None, num_classes=1000):
        super().__init__()
        if pretrained:
            raise RuntimeError("no network: torchvision ImageNet weights cannot be downloaded here; load a state_dict "
                               "(the key layout is torchvision's) after construction")
        self.inplanes = 64
        self.precision = precision or default_precision()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        self.layer4 = self._make_layer(512, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * Bottleneck.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.num_features = 512 * Bottleneck.expansion
        self._packed = None
        self._packed_key = None
        self.taps = None

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * Bottleneck.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * Bottleneck.expansion, kernel_size=1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes * Bottleneck.expansion))
        layers = [Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * Bottleneck.expansion
        layers += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    # ---------------------------------------------------------------- packed weights
    STEM_KPAD = 160   # 3*7*7 = 147 rounded up to the GEMM's K granule

    def _fingerprint(self):
        ts = list(self.parameters()) + [b for n, b in self.named_buffers() if "running" in n]
        return (self.precision, self.conv1.weight.device, sum(t._version for t in ts), id(self.conv1.weight))

    def packed(self):
        key = self._fingerprint()
        if self._packed is None or self._packed_key != key:
            wd = act_dtype(self.precision)

            def conv_bn(conv, bn):
                w, b = _fold_bn(conv, bn)
                n = w.shape[0]
                return w.permute(0, 2, 3, 1).reshape(n, -1).to(wd).contiguous(), b     # tap-major [N, R*S*C]

            w, b = _fold_bn(self.conv1, self.bn1)
            stem = torch.zeros(w.shape[0], self.STEM_KPAD, device=w.device, dtype=torch.float32)
            stem[:, :w[0].numel()] = w.reshape(w.shape[0], -1)                       # (c, ky, kx) order, zero padded
            pk = {"stem_w": stem.to(wd).contiguous(), "stem_b": b, "blocks": []}
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                for blk in layer:
                    d = dict(stride=blk.stride)
                    d["w1"], d["b1"] = conv_bn(blk.conv1, blk.bn1)
                    d["w2"], d["b2"] = conv_bn(blk.conv2, blk.bn2)
                    d["w3"], d["b3"] = conv_bn(blk.conv3, blk.bn3)
                    if blk.downsample is not None:
                        d["wd"], d["bd"] = conv_bn(blk.downsample[0], blk.downsample[1])
                    pk["blocks"].append(d)
            self._packed, self._packed_key = pk, key
        return self._packed

    # ---------------------------------------------------------------- forward
    def forward_features(self, x, final_gelu: bool = False):
        """-> NHWC matrix [B*Ho*Wo, 2048] in the activation dtype (+ the nn.GELU of model.py:232-235 when final_gelu)."""
        if self.training:
            raise NotImplementedError("train-mode BatchNorm (batch statistics) is outside the accelerated forward path; "
                                      "call model.eval()")
        if not x.is_cuda:
            raise RuntimeError("mvlt_b200 ResNet runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        B, Cin, H, W = x.shape
        assert Cin == 3, "ResNet stem expects 3 input channels"
        pk = self.packed()
        adt = act_dtype(self.precision)
        x = x.contiguous().float()
        if self.precision == "bf16":     # conv1 + bn1 + relu + maxpool as one tensor-core kernel
            a, H, W = ops.resnet_stem(x, pk["stem_w"], pk["stem_b"])
        else:                            # parity mode: patch matrix + CUDA-core GEMM + max-pool kernel
            a = ops.stem_im2col(x, 7, 7, 2, 3, self.STEM_KPAD, adt)
            a = ops.linear(a, pk["stem_w"], pk["stem_b"], act=ops.ACT_RELU)
            H, W = ops.conv_out_hw(H, W, 7, 7, 2, 3)
            a = ops.maxpool_nhwc(a, B, H, W, 3, 2, 1)
            H, W = ops.conv_out_hw(H, W, 3, 3, 2, 1)
        taps = self.taps
        if taps is not None:
            taps["stem"] = a.float().view(B, H, W, -1).permute(0, 3, 1, 2).clone()
        n_blocks = len(pk["blocks"])
        bi = 0
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4)):
            for _ in layer:
                w = pk["blocks"][bi]
                bi += 1
                s = w["stride"]
                h = ops.linear(a, w["w1"], w["b1"], act=ops.ACT_RELU)
                h = ops.conv2d_nhwc(h, w["w2"], w["b2"], B, H, W, 3, 3, s, 1, act=ops.ACT_RELU)
                idn = ops.conv2d_nhwc(a, w["wd"], w["bd"], B, H, W, 1, 1, s, 0) if "wd" in w else a
                H, W = ops.conv_out_hw(H, W, 3, 3, s, 1)
                last = final_gelu and bi == n_blocks
                a = ops.linear(h, w["w3"], w["b3"], act=ops.ACT_RELU_GELU if last else ops.ACT_RELU, residual=idn)
            if taps is not None and not (final_gelu and bi == n_blocks):   # layer4 of the fused-GELU call is not the raw output
                taps[f"layer{li + 1}"] = a.float().view(B, H, W, -1).permute(0, 3, 1, 2).clone()
        return a, H, W

    def forward(self, x):
        a, H, W = self.forward_features(x)
        return a.float().view(x.shape[0], H, W, -1).permute(0, 3, 1, 2)


class resnet101_without_fc(ResNetWithoutFC):
    """vfe.py:7-24."""

    def __init__(self, pretrained=False, progress=False, precision=None):
        super().__init__([3, 4, 23, 3], pretrained, progress, precision)


class resnet50_without_poolfc(ResNetWithoutFC):
    """vfe.py:27-44."""

    def __init__(self, pretrained=False, progress=False, precision=None):
        super().__init__([3, 4, 6, 3], pretrained, progress, precision)


# ====================================================================================================================
# Linear-patch and ViT backbones (vfe.py:47-107): 196 image tokens of width 768 (no resnet_fc), joint sequence 198 + L.
# ====================================================================================================================
class linear_patch_16x16(nn.Module):
    """vfe.py:47-60: Conv2d(3, 768, k=16, s=16) + BatchNorm2d(768) + ReLU -> [B, 768, 14, 14].  Non-overlapping patches: the
    convolution is a GEMM on the [B*196, 768] patch matrix (k = (c, ky, kx) = linear_patch.weight.view(768, -1) order) with the
    eval-mode BatchNorm folded into weight and bias and the ReLU (+ the GELU of model.py:232-235) in the epilogue."""

    def __init__(self, precision=None):
        super().__init__()
        self.linear_patch = nn.Conv2d(in_channels=3, out_channels=768, kernel_size=16, stride=16)
        self.bn = nn.BatchNorm2d(768)
        self.relu = nn.ReLU(inplace=True)
        self.precision = precision or default_precision()
        self._pk = None
        self.taps = None

    def packed(self):
        ts = [self.linear_patch.weight, self.linear_patch.bias, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        key = (self.precision, ts[0].device, sum(t._version for t in ts), id(ts[0]))
        if self._pk is None or self._pk[0] != key:
            w, b = _fold_bn(self.linear_patch, self.bn)
            self._pk = (key, w.reshape(w.shape[0], -1).to(act_dtype(self.precision)).contiguous(), b)
        return self._pk[1], self._pk[2]

    def forward_features(self, x, final_gelu: bool = False):
        """-> NHWC matrix [B*196, 768] in the activation dtype."""
        if self.training:
            raise NotImplementedError("train-mode BatchNorm (batch statistics) is outside the accelerated forward path; call model.eval()")
        if not x.is_cuda:
            raise RuntimeError("mvlt_b200 backbones run on CUDA (sm_100a) only; move the model and inputs to the GPU")
        w, b = self.packed()
        B = x.shape[0]
        a = ops.stem_im2col(x.contiguous().float(), 16, 16, 16, 0, w.shape[1], act_dtype(self.precision))
        a = ops.linear(a, w, b, act=ops.ACT_RELU_GELU if final_gelu else ops.ACT_RELU)
        H, W = x.shape[2] // 16, x.shape[3] // 16
        if self.taps is not None and not final_gelu:
            self.taps["linear_patch"] = a.float().view(B, H, W, -1).permute(0, 3, 1, 2).clone()
        return a, H, W

    def forward(self, x):
        a, H, W = self.forward_features(x)
        return a.float().view(x.shape[0], H, W, -1).permute(0, 3, 1, 2)


class _ViTMLP(nn.Sequential):
    """torchvision MLPBlock key layout: `0` Linear(768, 3072), `1` GELU, `2` Dropout, `3` Linear(3072, 768), `4` Dropout."""

    def __init__(self, d, hidden):
        super().__init__(nn.Linear(d, hidden), nn.GELU(), nn.Dropout(0.0), nn.Linear(hidden, d), nn.Dropout(0.0))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # torchvision < 0.13 (the reference pins >= 0.12, README.md:5) stored the block as `linear_1` / `linear_2`;
        # torchvision's own MLPBlock._load_from_state_dict performs the same rename
        for old, new in (("linear_1", "0"), ("linear_2", "3")):
            for kind in ("weight", "bias"):
                k = f"{prefix}{old}.{kind}"
                if k in state_dict:
                    state_dict[f"{prefix}{new}.{kind}"] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class _ViTEncoderBlock(nn.Module):
    def __init__(self, d, heads, hidden):
        super().__init__()
        self.num_heads = heads
        self.ln_1 = nn.LayerNorm(d, eps=1e-6)
        self.self_attention = nn.MultiheadAttention(d, heads, dropout=0.0, batch_first=True)   # in_proj_weight = Q|K|V packed
        self.dropout = nn.Dropout(0.0)
        self.ln_2 = nn.LayerNorm(d, eps=1e-6)
        self.mlp = _ViTMLP(d, hidden)


class _ViTEncoder(nn.Module):
    def __init__(self, seq, layers, heads, d, hidden):
        super().__init__()
        self.pos_embedding = nn.Parameter(torch.empty(1, seq, d).normal_(std=0.02))
        self.dropout = nn.Dropout(0.0)
        self.layers = nn.Sequential(collections.OrderedDict(
            (f"encoder_layer_{i}", _ViTEncoderBlock(d, heads, hidden)) for i in range(layers)))
        self.ln = nn.LayerNorm(d, eps=1e-6)


class VisionTransformerBaseWithoutPooling(nn.Module):
    """vfe.py:66-107: torchvision ViT-B/16 (`vit_b_16` key layout: class_token, conv_proj, encoder.pos_embedding,
    encoder.layers.encoder_layer_{i}.{ln_1, self_attention.{in_proj_weight, in_proj_bias, out_proj}, ln_2, mlp.{0,3}}, encoder.ln,
    heads.head) whose forward returns the 196 patch tokens `x[:, 1:]` without the classifier.  Pre-LN blocks on the same
    kernels as the BERT encoder: LayerNorm -> packed-QKV GEMM -> joint attention (197 tokens, no mask) -> out_proj GEMM
    accumulating in place into the fp32 stream -> LayerNorm -> fc1 + GELU -> fc2 in place; final LayerNorm."""

    def __init__(self, pretrained=False, progress=False, precision=None, **kwargs):
        super().__init__()
        if pretrained:
            raise RuntimeError("no network: torchvision ImageNet weights cannot be downloaded here; load a state_dict "
                               "(the key layout is torchvision's vit_b_16) after construction")
        d, heads, hidden, layers, patch, img = 768, 12, 3072, 12, 16, 224
        self.image_size, self.patch_size, self.hidden_dim, self.num_heads = img, patch, d, heads
        self.precision = precision or default_precision()
        self.class_token = nn.Parameter(torch.zeros(1, 1, d))
        self.conv_proj = nn.Conv2d(3, d, kernel_size=patch, stride=patch)
        self.seq_length = (img // patch) ** 2 + 1
        self.encoder = _ViTEncoder(self.seq_length, layers, heads, d, hidden)
        self.heads = nn.Sequential(collections.OrderedDict(head=nn.Linear(d, 1000)))      # constructed, never applied (vfe.py:101-105)
        fan_in = 3 * patch * patch
        nn.init.trunc_normal_(self.conv_proj.weight, std=(1 / fan_in) ** 0.5)
        nn.init.zeros_(self.conv_proj.bias)
        self._pk = None
        self.taps = None

    def packed(self):
        ps = list(self.parameters())
        key = (self.precision, ps[0].device, sum(p._version for p in ps), id(ps[0]))
        if self._pk is None or self._pk[0] != key:
            wd = act_dtype(self.precision)
            f32 = lambda t: t.detach().float().contiguous()
            wcast = lambda t: t.detach().to(wd).contiguous()
            pk = dict(proj_w=wcast(self.conv_proj.weight.reshape(self.hidden_dim, -1)), proj_b=f32(self.conv_proj.bias),
                      cls=f32(self.class_token.reshape(-1)), pos=f32(self.encoder.pos_embedding.reshape(self.seq_length, -1)),
                      ln_w=f32(self.encoder.ln.weight), ln_b=f32(self.encoder.ln.bias), layers=[])
            for blk in self.encoder.layers:
                at = blk.self_attention
                pk["layers"].append(dict(
                    n1w=f32(blk.ln_1.weight), n1b=f32(blk.ln_1.bias), n2w=f32(blk.ln_2.weight), n2b=f32(blk.ln_2.bias),
                    qkv_w=wcast(at.in_proj_weight), qkv_b=f32(at.in_proj_bias),
                    out_w=wcast(at.out_proj.weight), out_b=f32(at.out_proj.bias),
                    fc1_w=wcast(blk.mlp[0].weight), fc1_b=f32(blk.mlp[0].bias),
                    fc2_w=wcast(blk.mlp[3].weight), fc2_b=f32(blk.mlp[3].bias)))
            self._pk = (key, pk)
        return self._pk[1]

    def forward_features(self, x, final_gelu: bool = False):
        """-> fp32 [B, 196, 768]: the encoder output without the class token (+ the nn.GELU of model.py:232-235)."""
        if not x.is_cuda:
            raise RuntimeError("mvlt_b200 backbones run on CUDA (sm_100a) only; move the model and inputs to the GPU")
        B, _, H, W = x.shape
        assert H == self.image_size and W == self.image_size, f"Wrong image height/width! Expected {self.image_size} but got {H}x{W}!"
        pk = self.packed()
        adt = act_dtype(self.precision)
        p = self.patch_size
        a = ops.stem_im2col(x.contiguous().float(), p, p, p, 0, 3 * p * p, adt)
        patches = ops.linear(a, pk["proj_w"], pk["proj_b"], out_dtype=torch.float32)
        X = ops.vit_embed(patches, pk["cls"], pk["pos"], B)                       # fp32 [B*197, 768]
        S = self.seq_length
        kmask = torch.zeros((B, S), device=x.device, dtype=torch.float32)
        for li, (blk, w) in enumerate(zip(self.encoder.layers, pk["layers"])):
            h = ops.layernorm(X, w["n1w"], w["n1b"], blk.ln_1.eps, adt)
            qkv = ops.linear(h, w["qkv_w"], w["qkv_b"])
            ctx = ops.joint_attention(qkv, kmask, B, S, self.num_heads, False, S)
            ops.linear(ctx, w["out_w"], w["out_b"], residual=X, out=X)
            h = ops.layernorm(X, w["n2w"], w["n2b"], blk.ln_2.eps, adt)
            f = ops.linear(h, w["fc1_w"], w["fc1_b"], act=ops.ACT_GELU)
            ops.linear(f, w["fc2_w"], w["fc2_b"], residual=X, out=X)
            if self.taps is not None and li in (0, len(pk["layers"]) - 1):
                self.taps[f"vit{li}"] = X.clone().view(B, S, -1)
        out = ops.layernorm(X, pk["ln_w"], pk["ln_b"], self.encoder.ln.eps, torch.float32, gelu=final_gelu)
        return out.view(B, S, -1)[:, 1:]

    def forward(self, x):
        return self.forward_features(x)
