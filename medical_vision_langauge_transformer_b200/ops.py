"""Thin torch-tensor wrappers over the C ABI (raw `data_ptr()` + current CUDA stream).  PyTorch is used for device
memory and streams only; every arithmetic op below runs in libmvlt_b200.so.  No fallbacks."""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_TANH, ACT_RELU, ACT_RELU_GELU = 0, 1, 2, 3, 4
_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _rows2d(t: torch.Tensor):
    """(rows, cols, ld) of a tensor viewed as a row-major matrix with unit inner stride."""
    assert t.is_cuda and t.dim() >= 2 and t.stride(-1) == 1, "expected a CUDA tensor with contiguous last dim"
    if t.device.index != torch.cuda.current_device():
        # the C ABI launches on the CURRENT device's stream and keys its per-device setup on it (unlike a torch op, which
        # follows its tensors): refuse instead of launching on the wrong GPU with foreign pointers
        raise RuntimeError(f"mvlt_b200 ops run on the current CUDA device ({torch.cuda.current_device()}), but got a tensor on "
                           f"{t.device}; wrap the call in `with torch.cuda.device(tensor.device):`")
    if t.dim() == 2:
        return t.shape[0], t.shape[1], t.stride(0)
    assert t.is_contiguous()
    return t.numel() // t.shape[-1], t.shape[-1], t.shape[-1]


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           out_dtype: Optional[torch.dtype] = None, block_n: int = 0) -> torch.Tensor:
    """out = act(x @ w.T + bias) + residual (ACT_RELU / ACT_RELU_GELU: act applied after the residual add).
    bf16 x/w -> tcgen05 GEMM; fp32 x/w -> CUDA-core parity GEMM."""
    lib = _lib.ensure_init()
    M, K, lda = _rows2d(x)
    N, Kw, ldw = _rows2d(w)
    assert K == Kw, (x.shape, w.shape)
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=out_dtype or x.dtype)
    Mo, No, ldc = _rows2d(out)
    assert (Mo, No) == (M, N)
    ldres = 0
    if residual is not None:
        Mr, Nr, ldres = _rows2d(residual)
        assert (Mr, Nr) == (M, N)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    if x.dtype == torch.bfloat16:
        assert w.dtype == torch.bfloat16
        rc = lib.mvlt_gemm_bf16_tc(x.data_ptr(), lda, w.data_ptr(), ldw, out.data_ptr(), ldc, _ptr(bias), _ptr(residual),
                                   ldres, -1 if residual is None else _code(residual), M, N, K, act, _code(out), block_n,
                                   _stream())
        _lib.check(rc, f"mvlt_gemm_bf16_tc(M={M},N={N},K={K})")
    else:
        assert x.dtype == w.dtype == out.dtype == torch.float32
        assert residual is None or residual.dtype == torch.float32
        rc = lib.mvlt_gemm_f32_simt(x.data_ptr(), lda, w.data_ptr(), ldw, out.data_ptr(), ldc, _ptr(bias), _ptr(residual),
                                    ldres, M, N, K, act, _stream())
        _lib.check(rc, f"mvlt_gemm_f32_simt(M={M},N={N},K={K})")
    return out


def conv_out_hw(H: int, W: int, R: int, S: int, stride: int, pad: int):
    return (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1


def conv2d_nhwc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], B: int, H: int, W: int, R: int, S: int,
                stride: int, pad: int, act: int = ACT_NONE, residual: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, block_n: int = 0) -> torch.Tensor:
    """Conv2d (+ folded BatchNorm bias, + identity, + ReLU) over the NHWC activation x [B*H*W, C] -> [B*Ho*Wo, N].
    w [N, R*S*C] tap-major (k = (ky*S + kx)*C + c).  bf16: implicit GEMM, A fetched by im2col-mode TMA inside the tcgen05
    kernel; fp32 (parity mode): explicit patch matrix + CUDA-core GEMM."""
    lib = _lib.ensure_init()
    rows, C, ld = _rows2d(x)
    assert rows == B * H * W and ld == C and x.is_contiguous(), "NHWC activation must be a dense [B*H*W, C] matrix"
    N, K, ldw = _rows2d(w)
    assert K == R * S * C and w.dtype == x.dtype
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    M = B * Ho * Wo
    if x.dtype != torch.bfloat16:
        if R == S == 1 and stride == 1 and pad == 0:
            return linear(x, w, bias, act=act, residual=residual, out=out)
        patches = torch.empty((M, K), device=x.device, dtype=x.dtype)
        rc = lib.mvlt_im2col_nhwc(x.data_ptr(), _code(x), patches.data_ptr(), K, B, H, W, C, R, S, stride, pad, _stream())
        _lib.check(rc, "mvlt_im2col_nhwc")
        return linear(patches, w, bias, act=act, residual=residual, out=out)
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    Mo, No, ldc = _rows2d(out)
    assert (Mo, No) == (M, N) and out.dtype == torch.bfloat16
    ldres = 0
    if residual is not None:
        Mr, Nr, ldres = _rows2d(residual)
        assert (Mr, Nr) == (M, N) and residual.dtype == torch.bfloat16
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    rc = lib.mvlt_conv2d_nhwc_bf16_tc(x.data_ptr(), B, H, W, C, w.data_ptr(), ldw, out.data_ptr(), ldc, _ptr(bias),
                                      _ptr(residual), ldres, N, R, S, stride, pad, act, block_n, _stream())
    _lib.check(rc, f"mvlt_conv2d_nhwc_bf16_tc(B={B},H={H},W={W},C={C},N={N},k={R}x{S},s={stride})")
    return out


def stem_im2col(img: torch.Tensor, R: int, S: int, stride: int, pad: int, kpad: int, out_dtype: torch.dtype) -> torch.Tensor:
    """NCHW fp32 image -> patch matrix [B*Ho*Wo, kpad] (k = (c*R + ky)*S + kx, zero padded) for the ResNet stem conv."""
    lib = _lib.ensure_init()
    assert img.dtype == torch.float32 and img.is_contiguous() and img.dim() == 4
    B, Cin, H, W = img.shape
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    out = torch.empty((B * Ho * Wo, kpad), device=img.device, dtype=out_dtype)
    rc = lib.mvlt_stem_im2col_nchw(img.data_ptr(), out.data_ptr(), _code(out), kpad, B, Cin, H, W, R, S, stride, pad, kpad,
                                   _stream())
    _lib.check(rc, "mvlt_stem_im2col_nchw")
    return out


def resnet_stem(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor):
    """conv1 7x7/2 + folded BatchNorm + ReLU + MaxPool2d(3,2,1) in one kernel (bf16 mode).
    img fp32 NCHW [B,3,H,W]; w bf16 [64,160]; bias fp32 [64] -> (NHWC bf16 [B*Hp*Wp, 64], Hp, Wp)."""
    lib = _lib.ensure_init()
    assert img.dtype == torch.float32 and img.is_contiguous() and img.dim() == 4 and img.shape[1] == 3
    assert w.dtype == torch.bfloat16 and w.shape == (64, 160) and w.is_contiguous()
    assert bias.dtype == torch.float32 and bias.numel() == 64 and bias.is_contiguous()
    B, _, H, W = img.shape
    Ho, Wo = conv_out_hw(H, W, 7, 7, 2, 3)
    Hp, Wp = conv_out_hw(Ho, Wo, 3, 3, 2, 1)
    out = torch.empty((B * Hp * Wp, 64), device=img.device, dtype=torch.bfloat16)
    rc = lib.mvlt_resnet_stem_tc(img.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W, _stream())
    _lib.check(rc, "mvlt_resnet_stem_tc")
    return out, Hp, Wp


def maxpool_nhwc(x: torch.Tensor, B: int, H: int, W: int, k: int, stride: int, pad: int) -> torch.Tensor:
    lib = _lib.ensure_init()
    rows, C, ld = _rows2d(x)
    assert rows == B * H * W and ld == C and x.is_contiguous()
    Ho, Wo = conv_out_hw(H, W, k, k, stride, pad)
    out = torch.empty((B * Ho * Wo, C), device=x.device, dtype=x.dtype)
    rc = lib.mvlt_maxpool_nhwc(x.data_ptr(), out.data_ptr(), _code(x), B, H, W, C, k, stride, pad, _stream())
    _lib.check(rc, "mvlt_maxpool_nhwc")
    return out


FUSED_MLP_WIDTHS = (96, 192, 384)


def swin_mlp(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, w1: torch.Tensor, b1: torch.Tensor,
             w2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """x += fc2(gelu(fc1(layernorm(x)))) in place, one kernel (fp32 x, bf16 weights, C in FUSED_MLP_WIDTHS)."""
    lib = _lib.ensure_init()
    M, C, ldx = _rows2d(x)
    hidden = w1.shape[0]
    assert x.dtype == torch.float32 and w1.dtype == w2.dtype == torch.bfloat16
    assert w1.shape == (hidden, C) and w2.shape == (C, hidden) and w1.is_contiguous() and w2.is_contiguous()
    for t, n in ((gamma, C), (beta, C), (b1, hidden), (b2, C)):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    rc = lib.mvlt_swin_mlp_fused(x.data_ptr(), ldx, gamma.data_ptr(), beta.data_ptr(), float(eps), w1.data_ptr(),
                                 b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), M, C, hidden, _stream())
    _lib.check(rc, f"mvlt_swin_mlp_fused(M={M},C={C})")
    return x


LINEAR_LN_WIDTHS = (768,)

# ---- host-side routing between the one-wave fused kernels and the unfused chains --------------------------------------------
# The CTA-pair / cluster kernels own 256 rows per pair (cluster) for the whole contraction: they win when the row count fills most
# of a wave of the device's 74 SM pairs (33 four-CTA clusters) and lose to the finer-grained GEMM + row-kernel chain when it does
# not (tools/fused_crossover.py, profiles/r02_fused_crossover_by_batch.log: break-even near batch 40-48 of the bench workload).
# DEFAULT: the fused kernels always run — the forward stays a per-sample map BIT FOR BIT whatever the batch size (a slice of a
# batch reproduces its rows exactly, sharded retrieval equals single-rank retrieval; tests/test_e2e_gpu.py pins both).
# MVLT_ROUTING=auto picks per call by row count (faster below ~batch 40; results then depend on the batch size in the last bits).
_ROUTE = {}


def _route_info():
    dev = torch.cuda.current_device()
    if dev not in _ROUTE:
        lib = _lib.ensure_init()
        tiles = lib.mvlt_linear_ln_resident_tiles()
        if tiles <= 0:
            _lib.check(tiles if tiles < 0 else -2, "mvlt_linear_ln_resident_tiles")
        _ROUTE[dev] = (torch.cuda.get_device_properties(dev).multi_processor_count // 2, tiles)
    return _ROUTE[dev]


def _fills_wave(units: int, slots: int, min_first: float, min_eff: float) -> bool:
    if os.environ.get("MVLT_ROUTING", "fixed") != "auto":
        return True
    if units <= slots:
        return units >= min_first * slots
    waves = -(-units // slots)
    return units / (waves * slots) >= min_eff


def use_linear_ln(rows: int) -> bool:
    """linear_residual_layernorm vs GEMM (reduce-add epilogue) + layernorm: measured break-even at ~0.7 of a wave of clusters."""
    return _fills_wave(-(-rows // 256), _route_info()[1], 0.70, 0.85)


def use_block_tail(rows: int, C: int) -> bool:
    if C != 384:
        return True                     # C = 192: faster than the chain at every batch measured (8..64); C = 96 is opt-in anyway
    return _fills_wave(-(-rows // 256), _route_info()[0], 0.50, 0.85)


def use_ln_qkv(rows: int, C: int) -> bool:
    return _fills_wave(-(-rows // 256), _route_info()[0], 0.60, 0.90)



def linear_residual_layernorm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], residual: torch.Tensor, gamma: torch.Tensor,
                              beta: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None, bf16_copy: bool = True):
    """LayerNorm(a @ w.T + bias + residual) * gamma + beta in ONE kernel (bf16 a [M,K], bf16 w [N,K], fp32 residual [M,N], N in
    LINEAR_LN_WIDTHS) — the BertSelfOutput / BertOutput pattern.  -> fp32 [M,N] (`out`, may be `residual`), or (fp32, bf16 shadow)."""
    lib = _lib.ensure_init()
    M, K, lda = _rows2d(a)
    N = w.shape[0]
    Mr, Nr, ldr = _rows2d(residual)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.shape == (N, K) and w.is_contiguous()
    assert residual.dtype == torch.float32 and (Mr, Nr) == (M, N)
    for t in (gamma, beta) + ((bias,) if bias is not None else ()):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == N
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    Mo, No, ldo = _rows2d(out)
    assert out.dtype == torch.float32 and (Mo, No) == (M, N)
    shadow = torch.empty((M, N), device=a.device, dtype=torch.bfloat16) if bf16_copy else None
    rc = lib.mvlt_linear_residual_layernorm(a.data_ptr(), lda, w.data_ptr(), K, _ptr(bias), residual.data_ptr(), ldr, gamma.data_ptr(),
                                            beta.data_ptr(), float(eps), out.data_ptr(), ldo, _ptr(shadow), N, M, N, K, _stream())
    _lib.check(rc, f"mvlt_linear_residual_layernorm(M={M},N={N},K={K})")
    return (out, shadow) if bf16_copy else out


BLOCK_TAIL_WIDTHS = (96, 192, 384)
LN_QKV_WIDTHS = (96, 192, 384)


def swin_ln_qkv(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, w: torch.Tensor, bias: Optional[torch.Tensor],
                B: int, H: int, W: int, window: int, shift: int) -> torch.Tensor:
    """LayerNorm(x) @ w.T + bias with the OUTPUT rows window-major for the image rolled by -shift — norm1 + roll + window_partition +
    qkv of a Swin block in ONE kernel on CTA pairs (fp32 x [B*H*W, C] natural order, bf16 w [N, C], C in LN_QKV_WIDTHS)."""
    lib = _lib.ensure_init()
    M, C, ldx = _rows2d(x)
    N = w.shape[0]
    assert M == B * H * W and x.dtype == torch.float32 and w.dtype == torch.bfloat16 and w.shape == (N, C) and w.is_contiguous()
    for t, n in ((gamma, C), (beta, C)) + (((bias, N),) if bias is not None else ()):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    out = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    rc = lib.mvlt_swin_ln_qkv(x.data_ptr(), ldx, gamma.data_ptr(), beta.data_ptr(), float(eps), w.data_ptr(), _ptr(bias), out.data_ptr(),
                              B, H, W, C, N, window, shift, _stream())
    _lib.check(rc, f"mvlt_swin_ln_qkv(M={M},C={C},N={N})")
    return out



def swin_block_tail(x: torch.Tensor, o: Optional[torch.Tensor], wproj: Optional[torch.Tensor], bproj: Optional[torch.Tensor],
                    gamma: torch.Tensor, beta: torch.Tensor, eps: float, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
                    b2: torch.Tensor) -> torch.Tensor:
    """x += o @ wproj.T + bproj; x += fc2(gelu(fc1(layernorm(x)))) — in place, ONE kernel on CTA pairs (fp32 x [M, C], bf16 o and
    weights, C in BLOCK_TAIL_WIDTHS).  o=None: the MLP half only."""
    lib = _lib.ensure_init()
    M, C, ldx = _rows2d(x)
    hidden = w1.shape[0]
    assert x.dtype == torch.float32 and w1.dtype == w2.dtype == torch.bfloat16
    assert w1.shape == (hidden, C) and w2.shape == (C, hidden) and w1.is_contiguous() and w2.is_contiguous()
    if o is not None:
        assert o.dtype == torch.bfloat16 and o.shape == (M, C) and o.is_contiguous()
        assert wproj.dtype == torch.bfloat16 and wproj.shape == (C, C) and wproj.is_contiguous()
        assert bproj.dtype == torch.float32 and bproj.is_contiguous() and bproj.numel() == C
    for t, n in ((gamma, C), (beta, C), (b1, hidden), (b2, C)):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    rc = lib.mvlt_swin_block_tail(_ptr(o), x.data_ptr(), ldx, _ptr(wproj) if o is not None else None,
                                  _ptr(bproj) if o is not None else None, gamma.data_ptr(), beta.data_ptr(), float(eps),
                                  w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), M, C, hidden, _stream())
    _lib.check(rc, f"mvlt_swin_block_tail(M={M},C={C})")
    return x


FUSED_LN_LINEAR_WIDTHS = (192, 384)


def ln_linear(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, w: torch.Tensor,
              bias: Optional[torch.Tensor], act: int = ACT_NONE) -> torch.Tensor:
    """act(layernorm(x) @ w.T + bias) -> bf16 [M, N] in one kernel (fp32 x [M, C], bf16 w [N, C], C in FUSED_LN_LINEAR_WIDTHS)."""
    lib = _lib.ensure_init()
    M, C, ldx = _rows2d(x)
    N = w.shape[0]
    assert x.dtype == torch.float32 and w.dtype == torch.bfloat16 and w.shape == (N, C) and w.is_contiguous()
    for t, n in ((gamma, C), (beta, C)) + (((bias, N),) if bias is not None else ()):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    out = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    rc = lib.mvlt_ln_linear_bf16(x.data_ptr(), ldx, gamma.data_ptr(), beta.data_ptr(), float(eps), w.data_ptr(), _ptr(bias),
                                 out.data_ptr(), N, M, C, N, act, _stream())
    _lib.check(rc, f"mvlt_ln_linear_bf16(M={M},C={C},N={N})")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out_dtype: torch.dtype,
              gelu: bool = False, out: Optional[torch.Tensor] = None, bf16_copy: bool = False):
    """LayerNorm rows.  bf16_copy=True additionally returns the rows rounded to bf16 (-> (out, copy))."""
    lib = _lib.ensure_init()
    rows, C, ld = _rows2d(x)
    if out is None:
        out = torch.empty((rows, C), device=x.device, dtype=out_dtype)
    _, _, ldo = _rows2d(out)
    copy = torch.empty((rows, C), device=x.device, dtype=torch.bfloat16) if bf16_copy else None
    rc = lib.mvlt_layernorm_rows(x.data_ptr(), _code(x), ld, out.data_ptr(), _code(out), ldo, gamma.data_ptr(),
                                 beta.data_ptr(), rows, C, float(eps), int(gelu), _ptr(copy), C, _stream())
    _lib.check(rc, "mvlt_layernorm_rows")
    return (out, copy) if bf16_copy else out


def patch_embed_ln(img: torch.Tensor, w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                   eps: float = 1e-5, tensor_cores: bool = False, next_norm=None, next_norm_window: int = 0):
    """Conv2d(3,96,k4,s4) + LayerNorm(96).  tensor_cores=True: the mma.sync kernel with bf16 hi/lo operand splitting
    (fp32-accurate, the bf16-mode stem); False: the fp32 CUDA-core kernel (parity mode).
    next_norm=(gamma2, beta2, eps2) (tensor-core kernel only): also returns LayerNorm(out; gamma2, beta2) in bf16;
    next_norm_window > 0 writes that second output window-major (see layernorm_winmajor)."""
    lib = _lib.ensure_init()
    B = img.shape[0]
    assert img.dtype == torch.float32 and img.is_contiguous() and w.is_contiguous()
    E, P = w.shape[0], w.shape[-1]
    n_tok = (img.shape[-1] // P) ** 2
    out = torch.empty((B * n_tok, E), device=img.device, dtype=torch.float32)
    if tensor_cores:
        out2 = g2 = b2 = None
        eps2 = 0.0
        if next_norm is not None:
            g2, b2, eps2 = next_norm
            out2 = torch.empty((B * n_tok, E), device=img.device, dtype=torch.bfloat16)
        rc = lib.mvlt_patch_embed_ln_tc(img.data_ptr(), w.data_ptr(), b.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                        out.data_ptr(), B, img.shape[-1], P, E, float(eps), _ptr(g2), _ptr(b2), float(eps2),
                                        _ptr(out2), int(next_norm_window), _stream())
        _lib.check(rc, "mvlt_patch_embed_ln_tc")
        return (out, out2) if next_norm is not None else out
    assert next_norm is None
    rc = lib.mvlt_patch_embed_ln(img.data_ptr(), w.data_ptr(), b.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                 out.data_ptr(), B, img.shape[-1], P, E, float(eps), _stream())
    _lib.check(rc, "mvlt_patch_embed_ln")
    return out


def patch_merge_ln(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, B: int, H: int, W: int, C: int,
                   out_dtype: torch.dtype, eps: float = 1e-5) -> torch.Tensor:
    lib = _lib.ensure_init()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() == B * H * W * C
    out = torch.empty((B * (H // 2) * (W // 2), 4 * C), device=x.device, dtype=out_dtype)
    rc = lib.mvlt_patch_merge_ln(x.data_ptr(), out.data_ptr(), _code(out), gamma.data_ptr(), beta.data_ptr(), B, H, W, C,
                                 float(eps), _stream())
    _lib.check(rc, "mvlt_patch_merge_ln")
    return out


def layernorm_winmajor(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, B: int, H: int, W: int,
                       window: int, shift: int) -> torch.Tensor:
    """LayerNorm rows of the fp32 [B*H*W, C] token matrix -> bf16, written WINDOW-MAJOR for the image rolled by -shift:
    row (b*nW + w)*window^2 + i (vfe.py:356 + the roll / window_partition of :361-364)."""
    lib = _lib.ensure_init()
    rows, C, ld = _rows2d(x)
    assert x.dtype == torch.float32 and rows == B * H * W
    out = torch.empty((rows, C), device=x.device, dtype=torch.bfloat16)
    rc = lib.mvlt_layernorm_rows_winmajor(x.data_ptr(), ld, out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, H, W, C,
                                          window, shift, float(eps), _stream())
    _lib.check(rc, "mvlt_layernorm_rows_winmajor")
    return out


def window_major_index(B: int, H: int, W: int, window: int, shift: int, device=None) -> torch.Tensor:
    """perm[natural token row] = window-major row (the map of layernorm_winmajor); host-side helper for tests / tools."""
    r = torch.arange(B * H * W, device=device)
    b, rem = r // (H * W), r % (H * W)
    h, x = (rem // W - shift) % H, (rem % W - shift) % W
    nWw = W // window
    w, i = (h // window) * nWw + x // window, (h % window) * window + x % window
    return (b * (H // window) * nWw + w) * window * window + i


_LOG2E = 1.4426950408889634
WIN_BIAS_LD = 52


def window_bias_table(relbias: torch.Tensor, shift: int, window: int = 7) -> torch.Tensor:
    """Input-independent table of the tcgen05 window-attention kernel: fp32 [n_cls, heads, 49, 52] =
    (gathered relative-position bias [heads, 64, 64] (vfe.py:236-238) + the -100 shift mask (vfe.py:318-344)) * log2(e),
    columns 49..51 zero.  Window classes as in window_bias_fragments: bit 1 = last window row, bit 0 = last window column."""
    heads, n = relbias.shape[0], window * window
    dev = relbias.device
    i = torch.arange(n, device=dev)
    r, c = i // window, i % window
    n_cls = 4 if shift > 0 else 1
    out = torch.zeros(n_cls, heads, n, WIN_BIAS_LD, device=dev, dtype=torch.float32)
    for cls in range(n_cls):
        rh = torch.where(r < window - shift, 1, 2) if cls & 2 else torch.zeros_like(r)
        rw = torch.where(c < window - shift, 1, 2) if cls & 1 else torch.zeros_like(c)
        reg = rh * 3 + rw
        mask = torch.where(reg[:, None] != reg[None, :], -100.0, 0.0).to(torch.float32)
        out[cls, :, :, :n] = (relbias[:, :n, :n].float() + mask[None]) * _LOG2E
    return out.contiguous()


def window_attention_tc(qkv: torch.Tensor, bias_table: torch.Tensor, B: int, H: int, W: int, C: int, heads: int, window: int,
                        shift: int, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tcgen05 window attention: bf16 qkv [B*H*W, 3C] with WINDOW-MAJOR rows -> bf16 out [B*H*W, C] in natural token order."""
    lib = _lib.ensure_init()
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * H * W, 3 * C)
    assert bias_table.dtype == torch.float32 and bias_table.is_contiguous()
    assert bias_table.shape == (4 if shift > 0 else 1, heads, window * window, WIN_BIAS_LD), "takes window_bias_table()"
    if out is None:
        out = torch.empty((B * H * W, C), device=qkv.device, dtype=torch.bfloat16)
    rc = lib.mvlt_window_attention_tc(qkv.data_ptr(), out.data_ptr(), bias_table.data_ptr(), B, H, W, C, heads, window, shift,
                                      float(scale), _stream())
    _lib.check(rc, "mvlt_window_attention_tc")
    return out


def attention_impl() -> str:
    """'tc' (default): tcgen05 / TMEM / TMA attention kernels in bf16 mode; 'warp': the mma.sync kernels of round 1."""
    import os
    v = os.environ.get("MVLT_ATTN", "tc").strip().lower()
    if v not in ("tc", "warp"):
        raise ValueError(f"MVLT_ATTN={v!r}: expected 'tc' or 'warp'")
    return v


def joint_attention_tc_supported(S: int, head_dim: int) -> bool:
    return head_dim == 64 and (64 <= S <= 96 or 128 <= S <= 144)


_NEG_BIG = -1.0e30


def window_bias_fragments(relbias: torch.Tensor, shift: int, scale: float, window: int = 7) -> torch.Tensor:
    """Input-independent repacking of the gathered relative-position bias [heads, 64, 64] (vfe.py:236-238) for the bf16
    window-attention kernel: (bias + shift mask) / scale in the mma.m16n8 C-fragment order, one float4 per
    (window class, head, 16-row query tile, 8-key tile, lane); key columns 49..55 hold -1e30 (keys that do not exist).
    Window classes (shift > 0): bit 1 = last window row, bit 0 = last window column; the -100 mask of vfe.py:318-344
    separates the regions either side of the roll seam inside those windows.  shift == 0 -> one class."""
    heads, n = relbias.shape[0], window * window
    dev = relbias.device
    i = torch.arange(64, device=dev)
    r, c = i // window, i % window
    n_cls = 4 if shift > 0 else 1
    full = torch.zeros(n_cls, heads, 64, 64, device=dev, dtype=torch.float32)
    for cls in range(n_cls):
        rh = torch.where(r < window - shift, 1, 2) if cls & 2 else torch.zeros_like(r)
        rw = torch.where(c < window - shift, 1, 2) if cls & 1 else torch.zeros_like(c)
        reg = rh * 3 + rw
        mask = torch.where(reg[:, None] != reg[None, :], -100.0, 0.0).to(torch.float32)
        full[cls] = (relbias + mask[None]) / scale
    full[:, :, :, n:] = _NEG_BIG
    full[:, :, n:, :n] = 0.0
    lane = torch.arange(32, device=dev)
    g, t = lane // 4, lane % 4
    e = torch.arange(4, device=dev)
    mt, nt = torch.arange(4, device=dev), torch.arange(7, device=dev)
    rows = mt[:, None, None, None] * 16 + g[None, None, :, None] + (e[None, None, None, :] // 2) * 8      # [4,1,32,4]
    cols = nt[None, :, None, None] * 8 + 2 * t[None, None, :, None] + (e[None, None, None, :] % 2)        # [1,7,32,4]
    rows, cols = rows.expand(4, 7, 32, 4), cols.expand(4, 7, 32, 4)
    return full[:, :, rows, cols].contiguous()          # [n_cls, heads, 4, 7, 32, 4]


def window_attention(qkv: torch.Tensor, relbias: torch.Tensor, B: int, H: int, W: int, C: int, heads: int,
                     window: int, shift: int, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 qkv: relbias = gathered bias [heads, 64, 64].  bf16 qkv: relbias = window_bias_fragments(...) of it."""
    lib = _lib.ensure_init()
    assert qkv.is_contiguous() and qkv.shape == (B * H * W, 3 * C)
    assert relbias.dtype == torch.float32 and relbias.is_contiguous()
    if qkv.dtype == torch.bfloat16:
        assert relbias.shape == (4 if shift > 0 else 1, heads, 4, 7, 32, 4), "bf16 path takes window_bias_fragments()"
    else:
        assert relbias.shape == (heads, 64, 64)
    if out is None:
        out = torch.empty((B * H * W, C), device=qkv.device, dtype=qkv.dtype)
    rc = lib.mvlt_window_attention(qkv.data_ptr(), out.data_ptr(), _code(qkv), relbias.data_ptr(), B, H, W, C, heads,
                                   window, shift, float(scale), _stream())
    _lib.check(rc, "mvlt_window_attention")
    return out


def joint_embed(feat: torch.Tensor, ids: torch.Tensor, text_mask: Optional[torch.Tensor], image_mask: Optional[torch.Tensor],
                word_emb: torch.Tensor, typepos: torch.Tensor, cls_id: int, sep_id: int,
                img_index: Optional[torch.Tensor] = None, bf16_copy: bool = False):
    """-> (hidden [B*S, D] in feat.dtype, bf16 copy or None, kmask fp32 [B, S])"""
    lib = _lib.ensure_init()
    B, L = ids.shape
    n_obj, D = feat.shape[-2], feat.shape[-1]
    S = n_obj + 2 + L
    assert feat.is_contiguous() and ids.dtype == torch.int64 and ids.is_contiguous()
    assert typepos.shape[0] >= S and typepos.dtype == torch.float32 and word_emb.dtype == torch.float32
    tm = None if text_mask is None else text_mask.to(torch.uint8).contiguous()
    im = None if image_mask is None else image_mask.to(torch.uint8).contiguous()
    if img_index is not None:
        assert img_index.dtype == torch.int32 and img_index.numel() == B
    else:
        assert feat.shape[0] == B
    out = torch.empty((B * S, D), device=feat.device, dtype=feat.dtype)
    kmask = torch.empty((B, S), device=feat.device, dtype=torch.float32)
    copy = torch.empty((B * S, D), device=feat.device, dtype=torch.bfloat16) if bf16_copy else None
    rc = lib.mvlt_joint_embed(feat.data_ptr(), _code(feat), _ptr(img_index), ids.data_ptr(), _ptr(tm), _ptr(im),
                              word_emb.data_ptr(), typepos.data_ptr(), out.data_ptr(), _code(out), _ptr(copy),
                              kmask.data_ptr(), B, n_obj, L, D, int(cls_id), int(sep_id), int(word_emb.shape[0]), _stream())
    _lib.check(rc, "mvlt_joint_embed")
    return out, copy, kmask


def vit_embed(patches: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, B: int) -> torch.Tensor:
    """[B*n_patch, D] projected patches + class token + position embedding -> fp32 [B*(1+n_patch), D]."""
    lib = _lib.ensure_init()
    n_patch, D = patches.shape[0] // B, patches.shape[1]
    for t in (patches, cls, pos):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
    assert cls.numel() == D and pos.numel() == (n_patch + 1) * D
    out = torch.empty((B * (n_patch + 1), D), device=patches.device, dtype=torch.float32)
    rc = lib.mvlt_vit_embed(patches.data_ptr(), cls.data_ptr(), pos.data_ptr(), out.data_ptr(), B, n_patch, D, _stream())
    _lib.check(rc, "mvlt_vit_embed")
    return out


def joint_attention(qkv: torch.Tensor, kmask: torch.Tensor, B: int, S: int, heads: int, seq2seq: bool, obj_end: int,
                    out: Optional[torch.Tensor] = None, impl: Optional[str] = None) -> torch.Tensor:
    """bf16: the tcgen05 kernel for the joint sequence lengths it covers (impl 'tc', default), else the mma.sync kernel."""
    lib = _lib.ensure_init()
    C = qkv.shape[1] // 3
    hd = C // heads
    assert qkv.is_contiguous() and qkv.shape[0] == B * S and kmask.shape == (B, S)
    if out is None:
        out = torch.empty((B * S, C), device=qkv.device, dtype=qkv.dtype)
    if impl is None:
        impl = attention_impl()
    if qkv.dtype == torch.bfloat16 and impl == "tc" and joint_attention_tc_supported(S, hd):
        rc = lib.mvlt_joint_attention_tc(qkv.data_ptr(), out.data_ptr(), kmask.data_ptr(), B, S, heads, hd, int(seq2seq), obj_end,
                                         float(hd) ** -0.5, _stream())
        _lib.check(rc, "mvlt_joint_attention_tc")
        return out
    rc = lib.mvlt_joint_attention(qkv.data_ptr(), out.data_ptr(), _code(qkv), kmask.data_ptr(), B, S, heads, hd,
                                  int(seq2seq), obj_end, float(hd) ** -0.5, _stream())
    _lib.check(rc, "mvlt_joint_attention")
    return out


def linear_small(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    lib = _lib.ensure_init()
    rows, K, ld = _rows2d(x)
    N = w.shape[0]
    assert w.dtype == torch.float32 and w.is_contiguous() and w.shape[1] == K
    out = torch.empty((rows, N), device=x.device, dtype=torch.float32)
    rc = lib.mvlt_linear_small(x.data_ptr(), _code(x), ld, w.data_ptr(), _ptr(bias), out.data_ptr(), rows, N, K, _stream())
    _lib.check(rc, "mvlt_linear_small")
    return out


def softmax_rows(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.ensure_init()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    out = torch.empty_like(x)
    rc = lib.mvlt_softmax_rows(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _stream())
    _lib.check(rc, "mvlt_softmax_rows")
    return out


def masked_ce(logits: torch.Tensor, labels: torch.Tensor, n_classes: int, ignore_index: int = -100) -> torch.Tensor:
    """-> fp32 [2]: (sum of per-row losses over rows with label != ignore_index, number of such rows)."""
    lib = _lib.ensure_init()
    rows, _, ld = _rows2d(logits)
    assert logits.dtype == torch.float32 and labels.dtype == torch.int64 and labels.numel() == rows
    acc = torch.empty(2, device=logits.device, dtype=torch.float32)
    rc = lib.mvlt_masked_ce_rows(logits.data_ptr(), ld, labels.contiguous().data_ptr(), acc.data_ptr(), rows, n_classes,
                                 ignore_index, _stream())
    _lib.check(rc, "mvlt_masked_ce_rows")
    return acc


def mlm_ce_fused(t: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], labels: torch.Tensor,
                 ignore_index: int = -100) -> torch.Tensor:
    """Cross-entropy of the vocabulary projection t @ w.T + bias against labels WITHOUT materialising the logits (tcgen05 GEMM with
    an online-logsumexp epilogue).  t bf16 [rows, K], w bf16 [N, K] -> fp32 [2]: (sum of per-row losses over rows with
    label != ignore_index, number of such rows)."""
    lib = _lib.ensure_init()
    rows, K, ldt = _rows2d(t)
    N, Kw, ldw = _rows2d(w)
    assert t.dtype == w.dtype == torch.bfloat16 and K == Kw
    assert labels.dtype == torch.int64 and labels.numel() == rows
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    nbytes = lib.mvlt_mlm_ce_workspace_bytes(rows, N)
    ws = torch.empty(nbytes, device=t.device, dtype=torch.uint8)
    acc = torch.empty(2, device=t.device, dtype=torch.float32)
    rc = lib.mvlt_mlm_ce_fused(t.data_ptr(), ldt, w.data_ptr(), ldw, _ptr(bias), labels.contiguous().data_ptr(), acc.data_ptr(),
                               ws.data_ptr(), nbytes, rows, N, K, ignore_index, _stream())
    _lib.check(rc, f"mvlt_mlm_ce_fused(rows={rows},N={N},K={K})")
    return acc


def rank_first_positive(scores: torch.Tensor, labels: torch.Tensor):
    """compute_ranks of run_retrieval.py:220-249 on the device -> (row ranks int32 [R], column ranks int32 [C])."""
    lib = _lib.ensure_init()
    assert scores.is_cuda and scores.dtype == torch.float32 and scores.dim() == 2 and scores.stride(1) == 1
    lab = labels.to(device=scores.device).eq(1).to(torch.uint8).contiguous()
    R, C = scores.shape
    assert lab.shape == (R, C)
    rows = torch.empty(R, device=scores.device, dtype=torch.int32)
    cols = torch.empty(C, device=scores.device, dtype=torch.int32)
    rc = lib.mvlt_rank_first_positive(scores.data_ptr(), scores.stride(0), lab.data_ptr(), C, rows.data_ptr(), cols.data_ptr(),
                                      R, C, _stream())
    _lib.check(rc, "mvlt_rank_first_positive")
    return rows, cols


# ---- backward slice (csrc/backward.cu) ------------------------------------------------------------------------
def transpose_to_bf16(x: torch.Tensor, pad_to: int = 64) -> torch.Tensor:
    """[rows, cols] fp32 | bf16 -> bf16 [cols, rows rounded up to `pad_to`], zero padded: an M-contiguous wgrad operand."""
    lib = _lib.ensure_init()
    rows, cols, ld = _rows2d(x)
    ldo = -(-rows // pad_to) * pad_to
    out = torch.empty((cols, ldo), device=x.device, dtype=torch.bfloat16)
    rc = lib.mvlt_transpose_to_bf16(x.data_ptr(), _code(x), ld, out.data_ptr(), ldo, rows, cols, _stream())
    _lib.check(rc, "mvlt_transpose_to_bf16")
    return out


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, eps: float, bf16_copy: bool = True):
    """Backward of layer_norm(x) * gamma + beta for dense fp32 rows -> (dx fp32, dx bf16 | None, dgamma, dbeta)."""
    lib = _lib.ensure_init()
    rows, C, ld = _rows2d(x)
    assert dy.shape == x.shape and dy.dtype == x.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous() and ld == C
    dx = torch.empty_like(x)
    dxb = torch.empty((rows, C), device=x.device, dtype=torch.bfloat16) if bf16_copy else None
    dg, db = torch.empty(C, device=x.device), torch.empty(C, device=x.device)
    ws = torch.empty(max(lib.mvlt_layernorm_bwd_workspace_bytes(rows, C), 16), device=x.device, dtype=torch.uint8)
    rc = lib.mvlt_layernorm_bwd_rows(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), float(eps), dx.data_ptr(), _ptr(dxb), dg.data_ptr(),
                                     db.data_ptr(), ws.data_ptr(), rows, C, _stream())
    _lib.check(rc, f"mvlt_layernorm_bwd_rows(rows={rows},C={C})")
    return dx, dxb, dg, db


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of [rows, cols] (fp32 | bf16) -> fp32 [cols], fixed order."""
    lib = _lib.ensure_init()
    rows, cols, ld = _rows2d(x)
    out = torch.empty(cols, device=x.device)
    ws = torch.empty(max(lib.mvlt_colsum_workspace_bytes(rows, cols), 16), device=x.device, dtype=torch.uint8)
    rc = lib.mvlt_colsum(x.data_ptr(), _code(x), ld, out.data_ptr(), ws.data_ptr(), rows, cols, _stream())
    _lib.check(rc, "mvlt_colsum")
    return out


def gelu_bwd(u: torch.Tensor, df: torch.Tensor) -> torch.Tensor:
    """df * gelu'(u) (erf GELU), bf16 dense tensors of equal shape."""
    lib = _lib.ensure_init()
    assert u.shape == df.shape and u.dtype == df.dtype == torch.bfloat16 and u.is_contiguous() and df.is_contiguous()
    du = torch.empty_like(u)
    rc = lib.mvlt_gelu_bwd(u.data_ptr(), df.data_ptr(), du.data_ptr(), u.numel(), _stream())
    _lib.check(rc, "mvlt_gelu_bwd")
    return du


def joint_attention_bwd(qkv: torch.Tensor, kmask: Optional[torch.Tensor], dctx: torch.Tensor, B: int, S: int, heads: int, seq2seq: bool,
                        obj_end: int) -> torch.Tensor:
    """dqkv of joint_attention(qkv, kmask, ...) given dctx; bf16 [B*S, 3C] / [B*S, C] dense, head_dim 64."""
    lib = _lib.ensure_init()
    M, C3, ld = _rows2d(qkv)
    C = C3 // 3
    assert M == B * S and ld == C3 and qkv.dtype == dctx.dtype == torch.bfloat16 and dctx.shape == (M, C) and dctx.is_contiguous()
    dqkv = torch.empty_like(qkv)
    rc = lib.mvlt_joint_attention_bwd(qkv.data_ptr(), _ptr(kmask), dctx.data_ptr(), dqkv.data_ptr(), B, S, heads, C // heads, int(seq2seq),
                                      obj_end, float((C // heads) ** -0.5), _stream())
    _lib.check(rc, f"mvlt_joint_attention_bwd(B={B},S={S})")
    return dqkv
