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
        ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
        coords = torch.stack([ys.flatten(), xs.flatten()])
        d = coords[:, :, None] - coords[:, None, :]
        self.register_buffer("relative_position_index", (d[0] + wh - 1) * (2 * ww - 1) + (d[1] + ww - 1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        _trunc_normal_(self.relative_position_bias_table)

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
        mask = None
        if self.shift_size > 0:
            H, W = self.input_resolution
            ws, sh = self.window_size, self.shift_size
            region = torch.zeros(H, W)
            bounds = ((0, H - ws), (H - ws, H - sh), (H - sh, H))
            for a, (h0, h1) in enumerate(bounds):
                for b, (w0, w1) in enumerate(((0, W - ws), (W - ws, W - sh), (W - sh, W))):
                    region[h0:h1, w0:w1] = 3 * a + b
            win = region.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
            diff = win[:, None, :] - win[:, :, None]
            mask = torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))
        self.register_buffer("attn_mask", mask)


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
        dpr = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
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
                                 if self.precision == "bf16" else blk.attn.gathered_bias())))
                if layer.downsample is not None:
                    pk["merge"].append(dict(nw=f32(layer.downsample.norm.weight), nb=f32(layer.downsample.norm.bias),
                                            red_w=wcast(layer.downsample.reduction.weight)))
            self._packed, self._packed_key = pk, key
        return self._packed

    # ---------------------------------------------------------------- forward
    def forward_features(self, x, final_gelu: bool = False, out_dtype=None):
        pe = self.patch_embed
        B, C_in, H_img, W_img = x.shape
        assert H_img == pe.img_size[0] and W_img == pe.img_size[1], \
            f"Input image size ({H_img}*{W_img}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]})."
        if not x.is_cuda:
            raise RuntimeError("mvlt_b200 SwinTransformer runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        pk = self.packed()
        adt = act_dtype(self.precision)
        x = x.contiguous().float()
        a_first = None
        if self.precision == "bf16":   # tensor-core stem; norm1 of block 0 comes out of the same kernel
            b0 = self.layers[0].blocks[0]
            X, a_first = ops.patch_embed_ln(x, pk["pe_w"], pk["pe_b"], pe.norm.weight, pe.norm.bias, pe.norm.eps,
                                            tensor_cores=True, next_norm=(pk["blocks"][0]["n1w"], pk["blocks"][0]["n1b"], b0.norm1.eps))
        else:
            X = ops.patch_embed_ln(x, pk["pe_w"], pk["pe_b"], pe.norm.weight, pe.norm.bias, pe.norm.eps)
        # LN2 + fc1 + GELU + fc2 + residual as ONE tcgen05 kernel (csrc/swin_mlp.cu) for the stage widths listed in
        # MVLT_FUSED_MLP (bf16 mode).  Default: stage 0 only (C = 96, HBM-bound: 96 us vs 136 us for the unfused chain);
        # at C = 192 / 384 the kernel is shared-memory-bandwidth bound at its N = 64 MMAs and the unfused chain is faster
        # (profiles/r01_fused_mlp_vs_chain.log), "96,192,384" enables it everywhere, "" nowhere.
        fused_widths = tuple(int(v) for v in os.environ.get("MVLT_FUSED_MLP", "96").split(",") if v.strip()) \
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
                if a_first is not None:
                    a, a_first = a_first, None
                else:
                    a = ops.layernorm(X, w["n1w"], w["n1b"], blk.norm1.eps, adt)
                qkv = ops.linear(a, w["qkv_w"], w["qkv_b"])
                o = ops.window_attention(qkv, w["relbias"], B, H, W, C, blk.num_heads, blk.window_size, blk.shift_size,
                                         blk.attn.scale)
                ops.linear(o, w["proj_w"], w["proj_b"], residual=X, out=X)
                if C in fused_widths and C in ops.FUSED_MLP_WIDTHS and w["fc1_w"].shape[0] == 4 * C:
                    ops.swin_mlp(X, w["n2w"], w["n2b"], blk.norm2.eps, w["fc1_w"], w["fc1_b"], w["fc2_w"], w["fc2_b"])
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
