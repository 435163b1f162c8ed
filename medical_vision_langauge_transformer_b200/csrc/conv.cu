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

// ---------------------------------------------------------------------------------------------------------------------
// ResNet stem in one kernel (bf16 mode): conv1 7x7/2 pad 3 (3 -> 64) + folded BatchNorm + ReLU + MaxPool2d(3, 2, 1),
// NCHW fp32 image -> NHWC bf16 [B, Hp, Wp, 64]  (vfe.py:15-18).  A CTA iteration owns a 3 x 7 tile of POOLED pixels:
//   1. the 19 x 37 x 3 input patch behind it is staged in shared memory as bf16 (fetched into registers one tile ahead);
//   2. seven warps each compute one row of 16 convolution outputs x 64 channels on mma.sync m16n8k16.  K is laid out as
//      (c, ky, kx padded 7 -> 8) = 168, rounded up to 176: a k-pair (kx, kx+1) with kx even is ONE aligned 32-bit word of
//      the patch (input column 2*x + kx is even), so an A fragment register is a single LDS.32 through a per-thread table
//      of patch offsets held in registers; the weights are re-laid out to [64, 176] (zeros in the padding) while they are
//      staged in shared memory, where they stay for the whole kernel (ldmatrix);
//   3. bias + ReLU on the fragments, rows parked in shared memory (positions outside the image as 0 — equivalent to the
//      -inf padding of the max-pool because every window holds at least one real, non-negative value);
//   4. 3 x 3 / 2 max over the parked rows (bf16x2 max), 16-byte coalesced stores.
// The 7 x 16 convolution tile recomputes one halo row / column per pooled tile (1.33x the MMAs); nothing but the image is
// read from and nothing but the pooled map is written to HBM.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SP_ROWS = 3, SP_COLS = 7;                         // pooled tile
constexpr int SC_ROWS = 2 * SP_ROWS + 1, SC_COLS = 16;          // convolution tile (15 columns used)
constexpr int SI_ROWS = 2 * (SC_ROWS - 1) + 7, SI_COLS = 2 * (SC_COLS - 1) + 7, SI_LD = 40;   // 19 x 37 input patch
constexpr int SK = 3 * 7 * 7, SK_IN = 160;                      // filter elements; row length of the packed weights in HBM
constexpr int SK8 = 3 * 7 * 8, SKP = 176, SW_LD = 184, SO_LD = 72, STEM_N = 64;   // kx-padded K, smem row strides
constexpr int STEM_THREADS = 256;   // 8 warps (warps are allocated in fours: 9 would cost the registers of 12); 7 of them run MMAs
constexpr int STEM_SMEM = (STEM_N * SW_LD + 3 * SI_ROWS * SI_LD + SC_ROWS * SC_COLS * SO_LD) * 2 + STEM_N * 4;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// patch offset, in 32-bit words, of the k-pair starting at the (even) padded filter index k = (c*7 + ky)*8 + kx
__device__ __forceinline__ uint32_t stem_koff(int k) {
  if (k >= SK8) return 0;                                 // zero-weight padding: any finite value will do
  const int c = k / 56, r = k - c * 56, ky = r >> 3, kx = r & 7;
  return (uint32_t)(((c * SI_ROWS + ky) * SI_LD + kx) >> 1);
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

constexpr int STEM_PATCH_ELEMS = 3 * SI_ROWS * SI_COLS, STEM_PF = (STEM_PATCH_ELEMS + STEM_THREADS - 1) / STEM_THREADS;
// pe[j] packs (patch index << 16 | c << 12 | r << 6 | col) of the j-th patch element this thread fetches; 0xffffffff = none
__device__ __forceinline__ void stem_prefetch(float (&pf)[STEM_PF], const uint32_t (&pe)[STEM_PF], const float* __restrict__ img,
                                              int tile, int tiles_x, int tiles_y, int H, int W) {
  const int b = tile / (tiles_y * tiles_x), tr = tile - b * tiles_y * tiles_x;
  const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
  const int iy0 = 2 * (2 * ty * SP_ROWS - 1) - 3, ix0 = 2 * (2 * tx * SP_COLS - 1) - 3;
  const float* base = img + (long long)b * 3 * H * W;
#pragma unroll
  for (int j = 0; j < STEM_PF; ++j) {
    const int c = (pe[j] >> 12) & 3, iy = iy0 + (int)((pe[j] >> 6) & 63), ix = ix0 + (int)(pe[j] & 63);
    pf[j] = (pe[j] != 0xffffffffu && iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(base + ((long long)c * H + iy) * W + ix) : 0.f;
  }
}

__global__ void __launch_bounds__(STEM_THREADS, 2)
resnet_stem_tc_kernel(const float* __restrict__ img, const bf16* __restrict__ w, const float* __restrict__ bias,
                      bf16* __restrict__ out, int B, int H, int W, int Ho, int Wo, int Hp, int Wp) {
  extern __shared__ __align__(16) uint8_t smem_stem[];
  bf16* wsm = reinterpret_cast<bf16*>(smem_stem);                 // [64][SW_LD], k = (c*7 + ky)*8 + kx
  bf16* patch = wsm + STEM_N * SW_LD;                             // [3][SI_ROWS][SI_LD]
  bf16* stage = patch + 3 * SI_ROWS * SI_LD;                      // [SC_ROWS][SC_COLS][SO_LD]
  float* sbias = reinterpret_cast<float*>(stage + SC_ROWS * SC_COLS * SO_LD);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // word offsets of this thread's A-fragment k-pairs: k = 16*ks + 2t (+1) and 16*ks + 2t + 8 (+1)
  uint32_t kp[SKP / 16];
#pragma unroll
  for (int ks = 0; ks < SKP / 16; ++ks) kp[ks] = stem_koff(ks * 16 + 2 * t) | (stem_koff(ks * 16 + 2 * t + 8) << 16);
  uint32_t pe[STEM_PF];
#pragma unroll
  for (int j = 0; j < STEM_PF; ++j) {
    const int i = tid + j * STEM_THREADS;
    const int c = i / (SI_ROWS * SI_COLS), r2 = i - c * (SI_ROWS * SI_COLS);
    const int r = r2 / SI_COLS, col = r2 - r * SI_COLS;
    pe[j] = i < STEM_PATCH_ELEMS ? ((uint32_t)((c * SI_ROWS + r) * SI_LD + col) << 16) | (c << 12) | (r << 6) | col : 0xffffffffu;
  }
  // shared memory that is written once: zeroed weight tile (K padding), zero patch padding columns 37..39
  for (int i = tid; i < STEM_N * SW_LD / 8; i += STEM_THREADS) reinterpret_cast<uint4*>(wsm)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 3 * SI_ROWS * (SI_LD - SI_COLS); i += STEM_THREADS)
    patch[(i / (SI_LD - SI_COLS)) * SI_LD + SI_COLS + i % (SI_LD - SI_COLS)] = __float2bfloat16_rn(0.f);
  __syncthreads();
  pdl_grid_sync();
  for (int i = tid; i < STEM_N * SK; i += STEM_THREADS) {
    const int n = i / SK, k = i - n * SK;
    const int c = k / 49, r = k - c * 49, ky = r / 7, kx = r - ky * 7;
    wsm[n * SW_LD + (c * 7 + ky) * 8 + kx] = w[n * SK_IN + k];
  }
  if (tid < STEM_N) sbias[tid] = bias[tid];

  const int tiles_x = (Wp + SP_COLS - 1) / SP_COLS, tiles_y = (Hp + SP_ROWS - 1) / SP_ROWS;
  const int tiles = B * tiles_y * tiles_x;
  const uint32_t* patch32 = reinterpret_cast<const uint32_t*>(patch);
  // The input patch of the NEXT tile is fetched into registers while the current tile computes (all loads of a thread in
  // flight together), and written to shared memory at the top of the next iteration.
  float pf[STEM_PF];
  if (blockIdx.x < tiles) stem_prefetch(pf, pe, img, blockIdx.x, tiles_x, tiles_y, H, W);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b = tile / (tiles_y * tiles_x), tr = tile - b * tiles_y * tiles_x;
    const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
    const int py0 = ty * SP_ROWS, px0 = tx * SP_COLS;            // first pooled pixel
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;              // first convolution pixel (halo row / column)
    __syncthreads();                                              // previous iteration's pooling has read `stage`; weights landed
#pragma unroll
    for (int j = 0; j < STEM_PF; ++j)
      if (pe[j] != 0xffffffffu) patch[pe[j] >> 16] = __float2bfloat16_rn(pf[j]);
    __syncthreads();
    if (tile + (int)gridDim.x < tiles) stem_prefetch(pf, pe, img, tile + gridDim.x, tiles_x, tiles_y, H, W);

    // ---- convolution row `warp` of the tile: 16 pixels x 64 channels
    if (warp < SC_ROWS) {
    float acc[STEM_N / 8][4];
#pragma unroll
    for (int nt = 0; nt < STEM_N / 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    const uint32_t base0 = (uint32_t)(warp * SI_LD + g), base1 = base0 + 8;   // words: pixels g and g + 8 of the row
    const uint32_t wbase = smem_u32(wsm + ((lane >> 4) * 8 + (lane & 7)) * SW_LD + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int ks = 0; ks < SKP / 16; ++ks) {
      const uint32_t o0 = kp[ks] & 0xffffu, o1 = kp[ks] >> 16;
      const uint32_t a0 = patch32[base0 + o0], a1 = patch32[base1 + o0], a2 = patch32[base0 + o1], a3 = patch32[base1 + o1];
#pragma unroll
      for (int np = 0; np < STEM_N / 16; ++np) {
        uint32_t b0, b1, b2, b3;   // (n-tile 2np: k lo, k hi), (n-tile 2np+1: k lo, k hi)
        ldsm_x4(wbase + (uint32_t)((np * 16 * SW_LD + ks * 16) * 2), b0, b1, b2, b3);
        mma_bf16_16816(acc[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // ---- bias + ReLU, park the row (0 outside the image)
    const int cy = cy0 + warp;
    const bool row_ok = cy >= 0 && cy < Ho;
    const bool ok0 = row_ok && cx0 + g >= 0 && cx0 + g < Wo, ok1 = row_ok && cx0 + g + 8 >= 0 && cx0 + g + 8 < Wo;
    bf16* srow = stage + warp * SC_COLS * SO_LD;
#pragma unroll
    for (int nt = 0; nt < STEM_N / 8; ++nt) {
      const float2 bb = *reinterpret_cast<const float2*>(sbias + nt * 8 + 2 * t);
      const uint32_t v0 = ok0 ? pack_bf16x2(fmaxf(acc[nt][0] + bb.x, 0.f), fmaxf(acc[nt][1] + bb.y, 0.f)) : 0u;
      const uint32_t v1 = ok1 ? pack_bf16x2(fmaxf(acc[nt][2] + bb.x, 0.f), fmaxf(acc[nt][3] + bb.y, 0.f)) : 0u;
      *reinterpret_cast<uint32_t*>(srow + g * SO_LD + nt * 8 + 2 * t) = v0;
      *reinterpret_cast<uint32_t*>(srow + (g + 8) * SO_LD + nt * 8 + 2 * t) = v1;
    }
    }
    __syncthreads();
    // ---- 3x3/2 max-pool over the parked rows: one (pooled pixel, 8-channel chunk) per thread
    if (tid < SP_ROWS * SP_COLS * 8) {
      const int pp = tid >> 3, c8 = tid & 7;
      const int pr = pp / SP_COLS, pc = pp - pr * SP_COLS;
      const int py = py0 + pr, px = px0 + pc;
      if (py < Hp && px < Wp) {
        uint4 m = make_uint4(0, 0, 0, 0);   // values are >= 0
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint4 v = *reinterpret_cast<const uint4*>(stage + ((2 * pr + dy) * SC_COLS + 2 * pc + dx) * SO_LD + c8 * 8);
            m.x = bf16x2_max(m.x, v.x); m.y = bf16x2_max(m.y, v.y); m.z = bf16x2_max(m.z, v.z); m.w = bf16x2_max(m.w, v.w);
          }
        *reinterpret_cast<uint4*>(out + (((long long)b * Hp + py) * Wp + px) * STEM_N + c8 * 8) = m;
      }
    }
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

extern "C" int mvlt_resnet_stem_tc(const float* img, const void* w, const float* bias, void* out, int B, int H, int W,
                                   cudaStream_t stream) {
  if (!img || !w || !bias || !out || B <= 0 || H < 7 || W < 7) return MVLT_ERR_INVALID;
  if (((uintptr_t)w & 15) || ((uintptr_t)out & 15) || ((uintptr_t)bias & 7)) return MVLT_ERR_INVALID;
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;        // conv1: k 7, stride 2, pad 3
  const int Hp = (Ho + 2 - 3) / 2 + 1, Wp = (Wo + 2 - 3) / 2 + 1;      // maxpool: k 3, stride 2, pad 1
  const long long tiles = (long long)B * ((Hp + SP_ROWS - 1) / SP_ROWS) * ((Wp + SP_COLS - 1) / SP_COLS);
  if (tiles > 0x7fffffffLL) return MVLT_ERR_UNSUPPORTED;
  static unsigned long long attr_devices = 0;
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(resnet_stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (e != cudaSuccess) return (int)e;
    // two CTAs per SM
    e = cudaFuncSetAttribute(resnet_stem_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < 2LL * sms ? tiles : 2LL * sms);        // persistent: 2 CTAs per SM, weights staged once
  launch_k(resnet_stem_tc_kernel, dim3(grid), dim3(STEM_THREADS), STEM_SMEM, stream, img, (const bf16*)w, bias, (bf16*)out, B,
           H, W, Ho, Wo, Hp, Wp);
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
