// Memory-bound pieces of the ResNet visual backbone (reference: vfe.py:7-24 `resnet101_without_fc`, a torchvision
// ResNet with Bottleneck blocks; torchvision/models/resnet.py): activations live as NHWC matrices [B*H*W, C], every
// convolution is a GEMM over them (gemm_tc.cu: 1x1 convolutions directly, kxk / strided ones through im2col-mode TMA),
// eval-mode BatchNorm is folded into the packed weights and the epilogue bias by the host.
//
//   stem_im2col_kernel   NCHW fp32 image -> patch matrix [B*Ho*Wo, Kpad] for conv1 (7x7/2, 3 input channels: too narrow
//                        for a 64-channel im2col TMA box), K order (c, ky, kx) = conv1.weight.view(64, -1), zero padded
//   maxpool_nhwc_kernel  MaxPool2d(3, 2, 1) on NHWC
//   im2col_nhwc_kernel   explicit tap-major patch matrix [B*Ho*Wo, R*S*C] of an NHWC activation: the A operand of the
//                        fp32 parity-mode convolutions (gemm_simt.cu); the bf16 path never materialises it
#include "common.cuh"

namespace mvlt {

template <typename T> struct Vec16;  // 16-byte vector of T
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<bf16> { static constexpr int N = 8; };

template <typename T>
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const T* __restrict__ x, T* __restrict__ out, long long ld_out, int B, int H, int W, int C, int R, int S,
                   int stride, int pad, int Ho, int Wo) {
  pdl_grid_sync();
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;                                   // 16-byte chunks per pixel
  const long long per_row = (long long)R * S * cv;
  const long long total = (long long)B * Ho * Wo * per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / per_row;
    const int j = (int)(i - m * per_row);
    const int tap = j / cv, c = (j - tap * cv) * V;
    const int ky = tap / S, kx = tap - ky * S;
    const int n = (int)(m / (Ho * Wo));
    const int rem = (int)(m - (long long)n * Ho * Wo);
    const int p = rem / Wo, q = rem - p * Wo;
    const int h = p * stride - pad + ky, w = q * stride - pad + kx;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (h >= 0 && h < H && w >= 0 && w < W)
      v = *reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + c);
    *reinterpret_cast<uint4*>(out + m * ld_out + (long long)tap * C + c) = v;
  }
}

// two consecutive k per thread (kpad is even)
template <typename T>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, T* __restrict__ out, long long ld_out, int B, int Cin, int H, int W, int R,
                   int S, int stride, int pad, int Ho, int Wo, int kpad) {
  pdl_grid_sync();
  const int K = Cin * R * S, kp2 = kpad / 2;
  const long long total = (long long)B * Ho * Wo * kp2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / kp2;
    const int k0 = (int)(i - m * kp2) * 2;
    const int n = (int)(m / (Ho * Wo));
    const int rem = (int)(m - (long long)n * Ho * Wo);
    const int p = rem / Wo, q = rem - p * Wo;
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = k0 + e;
      v[e] = 0.f;
      if (k < K) {
        const int c = k / (R * S), t = k - c * R * S;
        const int ky = t / S, kx = t - ky * S;
        const int h = p * stride - pad + ky, w = q * stride - pad + kx;
        if (h >= 0 && h < H && w >= 0 && w < W) v[e] = __ldg(img + (((long long)n * Cin + c) * H + h) * W + w);
      }
    }
    T* o = out + m * ld_out + k0;
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v[0], v[1]);
    } else {
      *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool_nhwc_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int k, int stride, int pad,
                    int Ho, int Wo) {
  pdl_grid_sync();
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / cv;
    const int c = (int)(i - m * cv) * V;
    const int n = (int)(m / (Ho * Wo));
    const int rem = (int)(m - (long long)n * Ho * Wo);
    const int p = rem / Wo, q = rem - p * Wo;
    float best[V];
#pragma unroll
    for (int e = 0; e < V; ++e) best[e] = -INFINITY;   // nn.MaxPool2d pads with -inf
    for (int ky = 0; ky < k; ++ky) {
      const int h = p * stride - pad + ky;
      if (h < 0 || h >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int w = q * stride - pad + kx;
        if (w < 0 || w >= W) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + c);
        if constexpr (sizeof(T) == 2) {
          const uint32_t r[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(r[e]);
            best[2 * e] = fmaxf(best[2 * e], f.x);
            best[2 * e + 1] = fmaxf(best[2 * e + 1], f.y);
          }
        } else {
          best[0] = fmaxf(best[0], __uint_as_float(u.x));
          best[1] = fmaxf(best[1], __uint_as_float(u.y));
          best[2] = fmaxf(best[2], __uint_as_float(u.z));
          best[3] = fmaxf(best[3], __uint_as_float(u.w));
        }
      }
    }
    uint4 o;
    if constexpr (sizeof(T) == 2) {
      o = make_uint4(pack_bf16x2(best[0], best[1]), pack_bf16x2(best[2], best[3]), pack_bf16x2(best[4], best[5]),
                     pack_bf16x2(best[6], best[7]));
    } else {
      o = make_uint4(__float_as_uint(best[0]), __float_as_uint(best[1]), __float_as_uint(best[2]), __float_as_uint(best[3]));
    }
    *reinterpret_cast<uint4*>(out + m * (long long)C + c) = o;
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;   // grid-stride: 32 CTAs of 256 threads per SM is already more than resident
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mvlt

using namespace mvlt;

static bool conv_out_dims(int H, int W, int R, int S, int stride, int pad, int* Ho, int* Wo) {
  if (H <= 0 || W <= 0 || R <= 0 || S <= 0 || stride <= 0 || pad < 0) return false;
  *Ho = (H + 2 * pad - R) / stride + 1;
  *Wo = (W + 2 * pad - S) / stride + 1;
  return *Ho > 0 && *Wo > 0;
}

extern "C" int mvlt_im2col_nhwc(const void* x, int dtype, void* out, long long ld_out, int B, int H, int W, int C, int R,
                                int S, int stride, int pad, cudaStream_t stream) {
  int Ho, Wo;
  if (!x || !out || B <= 0 || C <= 0 || !conv_out_dims(H, W, R, S, stride, pad, &Ho, &Wo)) return MVLT_ERR_INVALID;
  if (((uintptr_t)x & 15) || ((uintptr_t)out & 15)) return MVLT_ERR_INVALID;
  const long long rows = (long long)B * Ho * Wo;
  if (dtype == MVLT_BF16) {
    if (C % 8 != 0 || ld_out % 8 != 0 || ld_out < (long long)R * S * C) return MVLT_ERR_INVALID;
    const long long total = rows * R * S * (C / 8);
    launch_k(im2col_nhwc_kernel<bf16>, dim3(grid_for(total, 256)), dim3(256), 0, stream, (const bf16*)x, (bf16*)out, ld_out, B,
             H, W, C, R, S, stride, pad, Ho, Wo);
  } else if (dtype == MVLT_F32) {
    if (C % 4 != 0 || ld_out % 4 != 0 || ld_out < (long long)R * S * C) return MVLT_ERR_INVALID;
    const long long total = rows * R * S * (C / 4);
    launch_k(im2col_nhwc_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, stream, (const float*)x, (float*)out, ld_out,
             B, H, W, C, R, S, stride, pad, Ho, Wo);
  } else {
    return MVLT_ERR_INVALID;
  }
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_stem_im2col_nchw(const float* img, void* out, int out_dtype, long long ld_out, int B, int Cin, int H,
                                     int W, int R, int S, int stride, int pad, int kpad, cudaStream_t stream) {
  int Ho, Wo;
  if (!img || !out || B <= 0 || Cin <= 0 || !conv_out_dims(H, W, R, S, stride, pad, &Ho, &Wo)) return MVLT_ERR_INVALID;
  if (kpad < Cin * R * S || kpad % 2 != 0 || ld_out < kpad || ((uintptr_t)out & 15)) return MVLT_ERR_INVALID;
  const long long total = (long long)B * Ho * Wo * (kpad / 2);
  if (out_dtype == MVLT_BF16) {
    if (ld_out % 2 != 0) return MVLT_ERR_INVALID;
    launch_k(stem_im2col_kernel<bf16>, dim3(grid_for(total, 256)), dim3(256), 0, stream, img, (bf16*)out, ld_out, B, Cin, H, W,
             R, S, stride, pad, Ho, Wo, kpad);
  } else if (out_dtype == MVLT_F32) {
    if (ld_out % 2 != 0) return MVLT_ERR_INVALID;
    launch_k(stem_im2col_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, stream, img, (float*)out, ld_out, B, Cin, H,
             W, R, S, stride, pad, Ho, Wo, kpad);
  } else {
    return MVLT_ERR_INVALID;
  }
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_maxpool_nhwc(const void* x, void* out, int dtype, int B, int H, int W, int C, int k, int stride, int pad,
                                 cudaStream_t stream) {
  int Ho, Wo;
  if (!x || !out || B <= 0 || C <= 0 || !conv_out_dims(H, W, k, k, stride, pad, &Ho, &Wo) || pad * 2 > k) return MVLT_ERR_INVALID;
  if (((uintptr_t)x & 15) || ((uintptr_t)out & 15)) return MVLT_ERR_INVALID;
  const long long rows = (long long)B * Ho * Wo;
  if (dtype == MVLT_BF16) {
    if (C % 8 != 0) return MVLT_ERR_INVALID;
    launch_k(maxpool_nhwc_kernel<bf16>, dim3(grid_for(rows * (C / 8), 256)), dim3(256), 0, stream, (const bf16*)x, (bf16*)out, B,
             H, W, C, k, stride, pad, Ho, Wo);
  } else if (dtype == MVLT_F32) {
    if (C % 4 != 0) return MVLT_ERR_INVALID;
    launch_k(maxpool_nhwc_kernel<float>, dim3(grid_for(rows * (C / 4), 256)), dim3(256), 0, stream, (const float*)x, (float*)out,
             B, H, W, C, k, stride, pad, Ho, Wo);
  } else {
    return MVLT_ERR_INVALID;
  }
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
