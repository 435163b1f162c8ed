"""Diagnostic for the im2col-mode TMA convolution: one-hot filter taps make the output a shifted copy of the input, so a
wrong corner / offset convention shows up as a wrong (dy, dx) instead of a bare mismatch.  python tools/conv_diag.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medical_vision_langauge_transformer_b200 import ops  # noqa: E402


def nhwc(x):
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous()


def main():
    torch.manual_seed(0)
    for (B, H, C, k, stride, pad) in [(2, 14, 64, 3, 1, 1), (2, 14, 64, 3, 2, 1), (2, 14, 64, 1, 2, 0), (3, 28, 128, 3, 1, 1)]:
        x = torch.randn(B, C, H, H, device="cuda").bfloat16()
        Ho = (H + 2 * pad - k) // stride + 1
        bad = 0
        for ky in range(k):
            for kx in range(k):
                w = torch.zeros(C, C, k, k, device="cuda")
                w[torch.arange(C), torch.arange(C), ky, kx] = 1.0
                wp = w.permute(0, 2, 3, 1).reshape(C, -1).bfloat16().contiguous()
                out = ops.conv2d_nhwc(nhwc(x), wp, None, B, H, H, k, k, stride, pad)
                torch.cuda.synchronize()
                ref = nhwc(F.conv2d(x.float(), w, stride=stride, padding=pad))
                err = (out.float() - ref).abs().max().item()
                if err > 1e-3:
                    bad += 1
                    found = None
                    o4 = out.float().view(B, Ho, Ho, C).permute(0, 3, 1, 2)
                    for dy in range(-4, 5):
                        for dx in range(-4, 5):
                            w2 = torch.zeros(C, C, 9, 9, device="cuda")
                            w2[torch.arange(C), torch.arange(C), dy + 4, dx + 4] = 1.0
                            cand = F.conv2d(x.float(), w2, stride=stride, padding=4)[:, :, :Ho, :Ho]
                            if cand.shape == o4.shape and (cand - o4).abs().max().item() < 1e-3:
                                found = (dy, dx)
                    print(f"  tap ({ky},{kx}) expected shift ({ky - pad},{kx - pad}) err {err:.3g} -> output matches shift {found}")
        print(f"B{B} H{H} C{C} k{k} s{stride} p{pad}: {'OK' if bad == 0 else f'{bad} taps wrong'}")
    # random weights, ragged sizes
    for (B, H, C, N, k, stride, pad) in [(5, 14, 256, 256, 3, 1, 1), (2, 56, 256, 512, 1, 2, 0), (7, 10, 128, 200, 3, 1, 1)]:
        x = torch.randn(B, C, H, H, device="cuda").bfloat16()
        w = (torch.randn(N, C, k, k, device="cuda") / (C * k * k) ** 0.5).bfloat16()
        out = ops.conv2d_nhwc(nhwc(x), w.permute(0, 2, 3, 1).reshape(N, -1).contiguous(), None, B, H, H, k, k, stride, pad)
        ref = nhwc(F.conv2d(x.float(), w.float(), stride=stride, padding=pad))
        print(f"B{B} H{H} C{C} N{N} k{k} s{stride}: relerr {((out.float() - ref).abs().max() / ref.abs().max()).item():.3g}")


if __name__ == "__main__":
    main()
