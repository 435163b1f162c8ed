// HBM-bound row kernels of the MVLT forward path: 128-bit vectorised loads/stores, one warp per row,
// warp-shuffle reductions, the row held in registers (read once, write once).
//
//   mvlt_layernorm_rows      vfe.py:356,:385 (norm1/norm2), HF modeling_bert.py:298,:356 (post-LN, eps 1e-12),
//                            vfe.py:685 + model.py:232-235 (final norm + GELU), HF :483 (head transform LN)
//   mvlt_patch_embed_ln      vfe.py:557-565 (Conv2d k4 s4 as a 48->96 dot + LayerNorm(96))
//   mvlt_patch_merge_ln      vfe.py:433-442 (2x2 gather in order (0,0),(1,0),(0,1),(1,1) + LayerNorm(4C))
//   mvlt_joint_embed         model.py:110-160 + :162-183 ([CLS] img [SEP] text + type + position; additive key mask)
#include "common.cuh"

namespace mvlt {

// Optional output-row permutation of the LayerNorm in front of a Swin qkv GEMM: natural token row (b, h, x) -> WINDOW-MAJOR
// row (b*nW + w)*ws*ws + i of the image rolled by -shift (vfe.py:361 torch.roll + :144-156 window_partition), so that the
// tcgen05 window-attention kernel fetches a window as one contiguous TMA box.  H == 0: identity.
struct WinMap {
  int H, W, ws, shift;
  __device__ __forceinline__ long long operator()(long long row) const {
    if (H == 0) return row;
    const int hw = H * W;
    const long long b = row / hw;
    const int rem = (int)(row - b * hw);
    int h = rem / W, x = rem - h * W;
    h -= shift; if (h < 0) h += H;        // the token at (h, x) sits at (h - shift, x - shift) of the rolled image
    x -= shift; if (x < 0) x += W;
    const int nWw = W / ws;
    const int w = (h / ws) * nWw + x / ws, i = (h % ws) * ws + x % ws;
    return (b * (long long)((H / ws) * nWw) + w) * (ws * ws) + i;
  }
};

// One warp normalises one row of C elements (C % 4 == 0, C <= 128*NCH).  `src(c)` returns the 4 elements
// starting at column c.  Two-pass statistics from registers (mean, then centred sum of squares).
template <int NCH, typename TO, typename SrcFn>
__device__ __forceinline__ void ln_row_core(SrcFn src, TO* __restrict__ dst, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, int C, float eps, bool gelu, int lane,
                                            bf16* __restrict__ dst_copy = nullptr) {
  float4 v[NCH];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < C) {
      v[i] = src(c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < C) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < C) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (gelu) {
        o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
      }
      store4(dst + c, o);
      if (dst_copy) store4(dst_copy + c, o);  // bf16 shadow = next GEMM's A operand; `dst` stays the fp32 residual
    }
  }
}

template <int NCH, typename TI, typename TO>
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const TI* __restrict__ in, long long ld_in, TO* __restrict__ out, long long ld_out,
                      const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int C, float eps,
                      int gelu, bf16* __restrict__ out_copy, long long ld_copy, const WinMap wm) {
  pdl_grid_sync();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* src = in + row * ld_in;
  ln_row_core<NCH>([&](int c) { return load4(src + c); }, out + wm(row) * ld_out, gamma, beta, C, eps, gelu != 0,
                   threadIdx.x & 31, out_copy ? out_copy + row * ld_copy : nullptr);
}

// Narrow rows (C % 32 == 0, C <= 128: Swin stage 0): eight lanes per row, four rows per warp, so every lane of the warp
// carries loads (one warp per 96-wide row leaves a quarter of the lanes idle and the kernel at half of HBM speed).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
layernorm_rows_small_kernel(const TI* __restrict__ in, long long ld_in, TO* __restrict__ out, long long ld_out,
                            const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int C, float eps,
                            int gelu, bf16* __restrict__ out_copy, long long ld_copy, const WinMap wm) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const long long row = (long long)blockIdx.x * 32 + (threadIdx.x >> 5) * 4 + (lane >> 3);
  const bool live = row < rows;
  const int nq = C >> 5;  // quads per lane
  const TI* src = in + (live ? row : 0) * ld_in;
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nq) {
      v[i] = load4(src + (sub + 8 * i) * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nq) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
  const float rstd = 1.0f / sqrtf(q / (float)C + eps);
  if (!live) return;
  const long long orow = wm(row);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nq) {
      const int c = (sub + 8 * i) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (gelu) {
        o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w);
      }
      store4(out + orow * ld_out + c, o);
      if (out_copy) store4(out_copy + row * ld_copy + c, o);
    }
}

// out row (b, h2, w2) = LN( cat[x(2h2,2w2), x(2h2+1,2w2), x(2h2,2w2+1), x(2h2+1,2w2+1)] ), x fp32 [B,H,W,C]
template <int NCH, typename TO>
__global__ void __launch_bounds__(256)
patch_merge_ln_kernel(const float* __restrict__ x, TO* __restrict__ out, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int B, int H, int W, int C, float eps) {
  pdl_grid_sync();
  const int H2 = H / 2, W2 = W / 2;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)B * H2 * W2) return;
  const int w2 = (int)(row % W2), h2 = (int)((row / W2) % H2), b = (int)(row / ((long long)W2 * H2));
  const float* base = x + (((long long)b * H + 2 * h2) * W + 2 * w2) * C;
  ln_row_core<NCH>(
      [&](int c) {
        const int qd = c / C, within = c - qd * C;  // quarter: dh = qd & 1, dw = qd >> 1
        return load4(base + ((long long)(qd & 1) * W + (qd >> 1)) * C + within);
      },
      out + row * 4LL * C, gamma, beta, 4 * C, eps, false, threadIdx.x & 31);
}

// One CTA per (image, patch row): 56 patches x 96 channels.  192 threads = 2 x 96: thread (half, c) keeps the 48
// weights of channel c in registers and walks every other patch; LayerNorm(96) by one warp per patch afterwards.
constexpr int PE_C = 96, PE_K = 48, PE_P = 56, PE_IMG = 224;
__global__ void __launch_bounds__(192)
patch_embed_ln_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                      float eps) {
  pdl_grid_sync();
  __shared__ __align__(16) float in_s[3][4][PE_IMG];
  __shared__ float out_s[PE_P][PE_C + 1];
  const int b = blockIdx.x / PE_P, py = blockIdx.x % PE_P;
  const int tid = threadIdx.x;
  for (int i = tid; i < 3 * 4 * (PE_IMG / 4); i += 192) {
    const int x4 = i % (PE_IMG / 4), ky = (i / (PE_IMG / 4)) % 4, ci = i / (4 * (PE_IMG / 4));
    const float4 t = load4(img + (((long long)b * 3 + ci) * PE_IMG + (4 * py + ky)) * PE_IMG + 4 * x4);
    *reinterpret_cast<float4*>(&in_s[ci][ky][4 * x4]) = t;
  }
  const int c = tid % PE_C, half = tid / PE_C;
  float wr[PE_K];
#pragma unroll
  for (int k = 0; k < PE_K; k += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(w + c * PE_K + k));
    wr[k] = t.x; wr[k + 1] = t.y; wr[k + 2] = t.z; wr[k + 3] = t.w;
  }
  const float bc = bias[c];
  __syncthreads();
  for (int p = half; p < PE_P; p += 2) {
    float acc = bc;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const float4 t = *reinterpret_cast<const float4*>(&in_s[ci][ky][4 * p]);
        acc = fmaf(wr[ci * 16 + ky * 4 + 0], t.x, acc);
        acc = fmaf(wr[ci * 16 + ky * 4 + 1], t.y, acc);
        acc = fmaf(wr[ci * 16 + ky * 4 + 2], t.z, acc);
        acc = fmaf(wr[ci * 16 + ky * 4 + 3], t.w, acc);
      }
    out_s[p][c] = acc;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int p = warp; p < PE_P; p += 6) {
    const float v0 = out_s[p][lane], v1 = out_s[p][lane + 32], v2 = out_s[p][lane + 64];
    const float mean = warp_sum(v0 + v1 + v2) * (1.0f / PE_C);
    const float d0 = v0 - mean, d1 = v1 - mean, d2 = v2 - mean;
    const float rstd = 1.0f / sqrtf(warp_sum(d0 * d0 + d1 * d1 + d2 * d2) * (1.0f / PE_C) + eps);
    float* o = out + ((long long)b * (PE_P * PE_P) + py * PE_P + p) * PE_C;
    o[lane] = d0 * rstd * gamma[lane] + beta[lane];
    o[lane + 32] = d1 * rstd * gamma[lane + 32] + beta[lane + 32];
    o[lane + 64] = d2 * rstd * gamma[lane + 64] + beta[lane + 64];
  }
}

// Tensor-core patch stem (bf16 mode): the same Conv2d(3,96,k4,s4) + LayerNorm(96) as a [B*3136, 48] x [48, 96] product
// on mma.sync m16n8k16, one warp per 16 consecutive patches.  The CUDA-core kernel above is FMA-bound (48 FMAs per
// output, 121 us at batch 64 against 18 us of HBM time for 38.5 MB in + 77 MB out); here the A fragments are read
// straight from the NCHW image (thread (g,t) of the fragment layout needs the float2 at image row 4py+ky, column
// 4px+2(t&1): every 32-byte sector it touches is fully used by its quad) and BOTH operands are split into bf16
// hi + lo halves (x = hi + lo to 16 mantissa bits) with all four cross products accumulated in fp32, so the result
// keeps fp32-level accuracy (the variance of RGC-shaped patch embeddings is ~eps: a plain bf16 stem would be a 1-2 %
// error after the LayerNorm).  Weights (hi, lo) sit in shared memory for ldmatrix; LayerNorm runs on the C fragments
// (quad shuffles), two-pass from registers.
constexpr int PT_WARPS = 4;
constexpr int PT_LDW = 56;  // bf16 elements per staged weight row (48 + 8: conflict-free ldmatrix)

__device__ __forceinline__ void pt_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void pt_ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void pt_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x -> (bf16x2 of the rounded values, bf16x2 of the rounding residuals)
__device__ __forceinline__ void pt_split(float2 x, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x.x, x.y);
  const float2 hf = __bfloat1622float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_bf16x2(x.x - hf.x, x.y - hf.y);
}

__global__ void __launch_bounds__(PT_WARPS * 32, 3)
patch_embed_ln_tc_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                         float eps, int n_groups, const float* __restrict__ gamma2, const float* __restrict__ beta2,
                         float eps2, bf16* __restrict__ out2, int out2_window) {
  __shared__ __align__(16) bf16 w_hi[PE_C * PT_LDW];
  __shared__ __align__(16) bf16 w_lo[PE_C * PT_LDW];
  __shared__ __align__(16) float4 s_gb[PE_C / 2];   // (gamma[c], gamma[c+1], beta[c], beta[c+1]) for even c
  __shared__ __align__(8) float2 s_bias[PE_C / 2];
  __shared__ __align__(16) float4 s_gb2[PE_C / 2];  // norm1 of the first Swin block (optional second output)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < PE_C * PE_K; i += PT_WARPS * 32) {   // parameters are static: staged before the PDL wait
    const int n = i / PE_K, k = i - n * PE_K;
    const float v = __ldg(w + i);
    const bf16 h = __float2bfloat16_rn(v);
    w_hi[n * PT_LDW + k] = h;
    w_lo[n * PT_LDW + k] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  for (int i = tid; i < PE_C / 2; i += PT_WARPS * 32) {
    s_gb[i] = make_float4(gamma[2 * i], gamma[2 * i + 1], beta[2 * i], beta[2 * i + 1]);
    s_bias[i] = make_float2(bias[2 * i], bias[2 * i + 1]);
    if (out2) s_gb2[i] = make_float4(gamma2[2 * i], gamma2[2 * i + 1], beta2[2 * i], beta2[2 * i + 1]);
  }
  __syncthreads();
  pdl_grid_sync();

  const uint32_t wh0 = smem_u32(w_hi), wl0 = smem_u32(w_lo);
  const uint32_t off4 = (uint32_t)(((lane & 7) * PT_LDW + (lane >> 3) * 8) * 2);              // k 0..31 of row lane&7
  const uint32_t off2 = (uint32_t)(((lane & 7) * PT_LDW + 32 + ((lane >> 3) & 1) * 8) * 2);   // k 32..47
  for (int grp = blockIdx.x * PT_WARPS + warp; grp < n_groups; grp += gridDim.x * PT_WARPS) {
    // rows g and g+8 of the 16-patch tile
    const float* src[2];
    float* dst[2];
    long long row2[2];   // row of the second output: natural, or window-major (window = out2_window, no shift) for the tcgen05 attention
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int P = grp * 16 + g + 8 * r;
      const int b = P / (PE_P * PE_P), rem = P - b * (PE_P * PE_P);
      const int py = rem / PE_P, px = rem - py * PE_P;
      row2[r] = P;
      if (out2_window > 0) {
        const int ws = out2_window, nWw = PE_P / ws;
        row2[r] = ((long long)b * (nWw * nWw) + (py / ws) * nWw + px / ws) * (ws * ws) + (py % ws) * ws + px % ws;
      }
      src[r] = img + ((long long)b * 3 * PE_IMG + 4 * py + (t >> 1)) * PE_IMG + 4 * px + 2 * (t & 1);
      dst[r] = out + (long long)P * PE_C + 2 * t;
    }
    float2 x[3][4];   // [ci][a0: row g ky t/2 | a1: row g+8 | a2: row g ky t/2+2 | a3: row g+8 ky t/2+2]
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int r = 0; r < 2; ++r)
          x[ci][2 * h + r] = __ldg(reinterpret_cast<const float2*>(src[r] + (long long)ci * PE_IMG * PE_IMG + 2 * h * PE_IMG));
    uint32_t ah[3][4], al[3][4];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int j = 0; j < 4; ++j) pt_split(x[ci][j], ah[ci][j], al[ci][j]);
    float acc[12][4];
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      const float2 bv = s_bias[nt * 4 + t];
      acc[nt][0] = bv.x; acc[nt][1] = bv.y; acc[nt][2] = bv.x; acc[nt][3] = bv.y;
      uint32_t bh[6], bl[6];
      const uint32_t row = (uint32_t)(nt * 8 * PT_LDW * 2);
      pt_ldsm_x4(wh0 + row + off4, bh[0], bh[1], bh[2], bh[3]);
      pt_ldsm_x2(wh0 + row + off2, bh[4], bh[5]);
      pt_ldsm_x4(wl0 + row + off4, bl[0], bl[1], bl[2], bl[3]);
      pt_ldsm_x2(wl0 + row + off2, bl[4], bl[5]);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        pt_mma(acc[nt], al[ci], bl[2 * ci], bl[2 * ci + 1]);   // smallest terms first
        pt_mma(acc[nt], al[ci], bh[2 * ci], bh[2 * ci + 1]);
        pt_mma(acc[nt], ah[ci], bl[2 * ci], bl[2 * ci + 1]);
        pt_mma(acc[nt], ah[ci], bh[2 * ci], bh[2 * ci + 1]);
      }
    }
    // LayerNorm(96) of rows g (acc[.][0..1]) and g+8 (acc[.][2..3]): the row is spread over the 4 lanes of a quad
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) { s0 += acc[nt][0] + acc[nt][1]; s1 += acc[nt][2] + acc[nt][3]; }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float m0 = s0 * (1.0f / PE_C), m1 = s1 * (1.0f / PE_C);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      acc[nt][0] -= m0; acc[nt][1] -= m0; acc[nt][2] -= m1; acc[nt][3] -= m1;
      q0 += acc[nt][0] * acc[nt][0] + acc[nt][1] * acc[nt][1];
      q1 += acc[nt][2] * acc[nt][2] + acc[nt][3] * acc[nt][3];
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float r0 = 1.0f / sqrtf(q0 * (1.0f / PE_C) + eps), r1 = 1.0f / sqrtf(q1 * (1.0f / PE_C) + eps);
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      const float4 gb = s_gb[nt * 4 + t];
      acc[nt][0] = acc[nt][0] * r0 * gb.x + gb.z; acc[nt][1] = acc[nt][1] * r0 * gb.y + gb.w;
      acc[nt][2] = acc[nt][2] * r1 * gb.x + gb.z; acc[nt][3] = acc[nt][3] * r1 * gb.y + gb.w;
      *reinterpret_cast<float2*>(dst[0] + nt * 8) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(dst[1] + nt * 8) = make_float2(acc[nt][2], acc[nt][3]);
    }
    if (out2) {  // norm1 of block 0 on the rows still in registers (vfe.py:356): the first LayerNorm launch of the trunk
      float u0 = 0.f, u1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) { u0 += acc[nt][0] + acc[nt][1]; u1 += acc[nt][2] + acc[nt][3]; }
      u0 += __shfl_xor_sync(0xffffffffu, u0, 1); u0 += __shfl_xor_sync(0xffffffffu, u0, 2);
      u1 += __shfl_xor_sync(0xffffffffu, u1, 1); u1 += __shfl_xor_sync(0xffffffffu, u1, 2);
      const float n0 = u0 * (1.0f / PE_C), n1 = u1 * (1.0f / PE_C);
      float p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) {
        acc[nt][0] -= n0; acc[nt][1] -= n0; acc[nt][2] -= n1; acc[nt][3] -= n1;
        p0 += acc[nt][0] * acc[nt][0] + acc[nt][1] * acc[nt][1];
        p1 += acc[nt][2] * acc[nt][2] + acc[nt][3] * acc[nt][3];
      }
      p0 += __shfl_xor_sync(0xffffffffu, p0, 1); p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
      p1 += __shfl_xor_sync(0xffffffffu, p1, 1); p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
      const float z0 = 1.0f / sqrtf(p0 * (1.0f / PE_C) + eps2), z1 = 1.0f / sqrtf(p1 * (1.0f / PE_C) + eps2);
      bf16* d0 = out2 + row2[0] * PE_C + 2 * t, * d1 = out2 + row2[1] * PE_C + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) {
        const float4 gb = s_gb2[nt * 4 + t];
        *reinterpret_cast<uint32_t*>(d0 + nt * 8) = pack_bf16x2(acc[nt][0] * z0 * gb.x + gb.z, acc[nt][1] * z0 * gb.y + gb.w);
        *reinterpret_cast<uint32_t*>(d1 + nt * 8) = pack_bf16x2(acc[nt][2] * z1 * gb.x + gb.z, acc[nt][3] * z1 * gb.y + gb.w);
      }
    }
  }
}

// One warp per joint-sequence row (b, s), D % 4 == 0, D <= 128*NCH.
//   s in [1, n_obj]      : image feature row (already LN+GELU'd)          model.py:141
//   s == 0 / n_obj+1     : word_emb[cls_id] / word_emb[sep_id]            model.py:133-136
//   s >  n_obj+1         : word_emb[ids[b, s-n_obj-2]]                    model.py:138
//   + typepos[s] = token_type_emb[s <= n_obj+1] + position_emb[s]         model.py:152-158 (precomputed table)
// kmask[b,s] = 0 where attended, -10000 where masked (model.py:126,:182).
template <int NCH, typename TF, typename TO>
__global__ void __launch_bounds__(256)
joint_embed_kernel(const TF* __restrict__ feat, const int* __restrict__ img_index, const long long* __restrict__ ids,
                   const unsigned char* __restrict__ text_mask, const unsigned char* __restrict__ image_mask,
                   const float* __restrict__ word_emb, const float* __restrict__ typepos, TO* __restrict__ out,
                   bf16* __restrict__ out_copy, float* __restrict__ kmask, int B, int n_obj, int L, int D, int cls_id,
                   int sep_id, int vocab_rows) {
  pdl_grid_sync();
  const int S = n_obj + 2 + L;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)B * S) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / S), s = (int)(row % S);
  const float* tp = typepos + (long long)s * D;
  TO* o = out + row * D;
  bf16* o2 = out_copy ? out_copy + row * D : nullptr;
  bool keep = true;
  if (s >= 1 && s <= n_obj) {
    const int fi = img_index ? img_index[b] : b;
    const TF* f = feat + ((long long)fi * n_obj + (s - 1)) * D;
    if (image_mask) keep = image_mask[(long long)b * n_obj + (s - 1)] != 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < D) {
        const float4 a = load4(f + c), t = __ldg(reinterpret_cast<const float4*>(tp + c));
        const float4 r = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
        store4(o + c, r);
        if (o2) store4(o2 + c, r);
      }
    }
  } else {
    long long id = cls_id;
    if (s == n_obj + 1) id = sep_id;
    else if (s > n_obj + 1) {
      id = ids[(long long)b * L + (s - n_obj - 2)];
      keep = text_mask ? text_mask[(long long)b * L + (s - n_obj - 2)] != 0 : id > 0;  // model.py:337 (ids > 0)
    }
    // nn.Embedding raises on an id outside the table; a kernel cannot: the row becomes NaN instead of an out-of-bounds read
    const bool id_ok = id >= 0 && id < vocab_rows;
    const float* e = word_emb + (id_ok ? id : 0) * D;
    const float poison = id_ok ? 0.f : __int_as_float(0x7fc00000);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < D) {
        float4 a = __ldg(reinterpret_cast<const float4*>(e + c));
        const float4 t = __ldg(reinterpret_cast<const float4*>(tp + c));
        a.x += poison;
        const float4 r = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
        store4(o + c, r);
        if (o2) store4(o2 + c, r);
      }
    }
  }
  if (lane == 0) kmask[row] = keep ? 0.0f : -10000.0f;
}

template <typename TI, typename TO>
static int launch_ln(const void* in, long long ld_in, void* out, long long ld_out, const float* gamma, const float* beta,
                     long long rows, int C, float eps, int gelu, void* out_copy, long long ld_copy, cudaStream_t st,
                     WinMap wm = WinMap{0, 0, 1, 0}) {
  if (C <= 128 && C % 32 == 0) {
    launch_k(layernorm_rows_small_kernel<TI, TO>, dim3((unsigned)((rows + 31) / 32)), dim3(256), 0, st, (const TI*)in, ld_in,
             (TO*)out, ld_out, gamma, beta, rows, C, eps, gelu, (bf16*)out_copy, ld_copy, wm);
    return MVLT_OK;
  }
  const unsigned grid = (unsigned)((rows + 7) / 8);
#define LN_CASE(NCH)                                                                                              \
  launch_k(layernorm_rows_kernel<NCH, TI, TO>, dim3(grid), dim3(256), 0, st, (const TI*)in, ld_in, (TO*)out, ld_out, gamma, beta, \
                                                           rows, C, eps, gelu, (bf16*)out_copy, ld_copy, wm)
  if (C <= 128) LN_CASE(1);
  else if (C <= 256) LN_CASE(2);
  else if (C <= 384) LN_CASE(3);
  else if (C <= 768) LN_CASE(6);
  else if (C <= 1536) LN_CASE(12);
  else if (C <= 3072) LN_CASE(24);
  else return MVLT_ERR_UNSUPPORTED;
#undef LN_CASE
  return MVLT_OK;
}

// ViT token assembly (torchvision vision_transformer.py: `_process_input` + class token + `pos_embedding`, as used by
// vfe.py:94-107): out[b, 0] = class_token + pos[0]; out[b, 1 + i] = patch_proj[b, i] + pos[1 + i].  One float4 per thread.
__global__ void __launch_bounds__(256)
vit_embed_kernel(const float* __restrict__ patches, const float* __restrict__ cls, const float* __restrict__ pos,
                 float* __restrict__ out, int B, int n_patch, int D) {
  pdl_grid_sync();
  const int dq = D >> 2, S = n_patch + 1;
  const long long total = (long long)B * S * dq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dq) * 4;
    const long long r = i / dq;
    const int s = (int)(r % S), b = (int)(r / S);
    const float4 a = s == 0 ? load4(cls + c) : load4(patches + ((long long)b * n_patch + s - 1) * D + c);
    const float4 q = load4(pos + (long long)s * D + c);
    store4(out + r * D + c, make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w));
  }
}

}  // namespace mvlt

using namespace mvlt;

extern "C" int mvlt_layernorm_rows(const void* in, int in_dtype, long long ld_in, void* out, int out_dtype,
                                   long long ld_out, const float* gamma, const float* beta, long long rows, int C,
                                   float eps, int gelu, void* out_bf16_copy, long long ld_copy, cudaStream_t stream) {
  if (!in || !out || !gamma || !beta || rows <= 0 || C <= 0 || C % 4 || ld_in % 4 || ld_out % 4) return MVLT_ERR_INVALID;
  if (out_bf16_copy && ld_copy % 4) return MVLT_ERR_INVALID;
  int rc;
  if (in_dtype == MVLT_F32 && out_dtype == MVLT_F32) rc = launch_ln<float, float>(in, ld_in, out, ld_out, gamma, beta, rows, C, eps, gelu, out_bf16_copy, ld_copy, stream);
  else if (in_dtype == MVLT_F32 && out_dtype == MVLT_BF16) rc = launch_ln<float, bf16>(in, ld_in, out, ld_out, gamma, beta, rows, C, eps, gelu, out_bf16_copy, ld_copy, stream);
  else if (in_dtype == MVLT_BF16 && out_dtype == MVLT_BF16) rc = launch_ln<bf16, bf16>(in, ld_in, out, ld_out, gamma, beta, rows, C, eps, gelu, out_bf16_copy, ld_copy, stream);
  else if (in_dtype == MVLT_BF16 && out_dtype == MVLT_F32) rc = launch_ln<bf16, float>(in, ld_in, out, ld_out, gamma, beta, rows, C, eps, gelu, out_bf16_copy, ld_copy, stream);
  else return MVLT_ERR_INVALID;
  if (rc != MVLT_OK) return rc;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

// LayerNorm(x) -> bf16 with the output rows in WINDOW-MAJOR order of the image rolled by -shift (see WinMap): norm1 of a Swin
// block (vfe.py:356) fused with torch.roll + window_partition (vfe.py:361-367).  x: fp32 [B*H*W, C] natural token order.
extern "C" int mvlt_layernorm_rows_winmajor(const float* in, long long ld_in, void* out_bf16, const float* gamma, const float* beta,
                                            int B, int H, int W, int C, int window, int shift, float eps, cudaStream_t stream) {
  if (!in || !out_bf16 || !gamma || !beta || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 4 || ld_in % 4) return MVLT_ERR_INVALID;
  if (window <= 0 || H % window || W % window || shift < 0 || shift >= window) return MVLT_ERR_INVALID;
  const int rc = launch_ln<float, bf16>(in, ld_in, out_bf16, C, gamma, beta, (long long)B * H * W, C, eps, 0, nullptr, 0, stream,
                                        WinMap{H, W, window, shift});
  if (rc != MVLT_OK) return rc;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_patch_embed_ln(const float* img, const float* weight, const float* bias, const float* gamma,
                                   const float* beta, float* out, int B, int img_size, int patch, int embed_dim,
                                   float eps, cudaStream_t stream) {
  if (!img || !weight || !bias || !gamma || !beta || !out || B <= 0) return MVLT_ERR_INVALID;
  if (img_size != PE_IMG || patch != 4 || embed_dim != PE_C) return MVLT_ERR_UNSUPPORTED;  // Swin-S/T/B-224 patch stem
  launch_k(patch_embed_ln_kernel, dim3(B * PE_P), dim3(192), 0, stream, img, weight, bias, gamma, beta, out, eps);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

// Same contract on the tensor cores (bf16-mode stem; fp32-level accuracy through bf16 hi/lo operand splitting).
extern "C" int mvlt_patch_embed_ln_tc(const float* img, const float* weight, const float* bias, const float* gamma,
                                      const float* beta, float* out, int B, int img_size, int patch, int embed_dim,
                                      float eps, const float* gamma2, const float* beta2, float eps2, void* out2_bf16,
                                      int out2_window, cudaStream_t stream) {
  if (!img || !weight || !bias || !gamma || !beta || !out || B <= 0) return MVLT_ERR_INVALID;
  if (out2_bf16 && (!gamma2 || !beta2 || ((uintptr_t)out2_bf16 & 3))) return MVLT_ERR_INVALID;
  if (img_size != PE_IMG || patch != 4 || embed_dim != PE_C) return MVLT_ERR_UNSUPPORTED;
  if (out2_window < 0 || (out2_window > 0 && PE_P % out2_window != 0)) return MVLT_ERR_INVALID;
  if (((uintptr_t)img & 7) || ((uintptr_t)out & 7)) return MVLT_ERR_INVALID;
  const int n_groups = B * (PE_P * PE_P / 16);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ctas = (n_groups + PT_WARPS - 1) / PT_WARPS;
  launch_k(patch_embed_ln_tc_kernel, dim3(ctas < 3 * sms ? ctas : 3 * sms), dim3(PT_WARPS * 32), 0, stream, img, weight, bias,
           gamma, beta, out, eps, n_groups, gamma2, beta2, eps2, reinterpret_cast<bf16*>(out2_bf16), out2_window);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_patch_merge_ln(const float* x, void* out, int out_dtype, const float* gamma, const float* beta,
                                   int B, int H, int W, int C, float eps, cudaStream_t stream) {
  if (!x || !out || !gamma || !beta || B <= 0 || H % 2 || W % 2 || C % 4) return MVLT_ERR_INVALID;
  const long long rows = (long long)B * (H / 2) * (W / 2);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const int C4 = 4 * C;
#define PM_CASE(NCH)                                                                                            \
  do {                                                                                                          \
    if (out_dtype == MVLT_F32) launch_k(patch_merge_ln_kernel<NCH, float>, dim3(grid), dim3(256), 0, stream, x, (float*)out, gamma, beta, B, H, W, C, eps); \
    else launch_k(patch_merge_ln_kernel<NCH, bf16>, dim3(grid), dim3(256), 0, stream, x, (bf16*)out, gamma, beta, B, H, W, C, eps); \
  } while (0)
  if (out_dtype != MVLT_F32 && out_dtype != MVLT_BF16) return MVLT_ERR_INVALID;
  if (C4 <= 384) PM_CASE(3);
  else if (C4 <= 768) PM_CASE(6);
  else if (C4 <= 1536) PM_CASE(12);
  else if (C4 <= 3072) PM_CASE(24);
  else return MVLT_ERR_UNSUPPORTED;
#undef PM_CASE
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_joint_embed(const void* feat, int feat_dtype, const int* img_index, const long long* ids,
                                const unsigned char* text_mask, const unsigned char* image_mask, const float* word_emb,
                                const float* typepos, void* out, int out_dtype, void* out_bf16_copy, float* kmask, int B,
                                int n_obj, int L, int D, int cls_id, int sep_id, int vocab_rows, cudaStream_t stream) {
  if (!feat || !ids || !word_emb || !typepos || !out || !kmask || B <= 0 || n_obj <= 0 || L < 0) return MVLT_ERR_INVALID;
  if (D != 768) return MVLT_ERR_UNSUPPORTED;
  if (feat_dtype != out_dtype || vocab_rows <= 0 || cls_id < 0 || cls_id >= vocab_rows || sep_id < 0 || sep_id >= vocab_rows) return MVLT_ERR_INVALID;
  const long long rows = (long long)B * (n_obj + 2 + L);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (out_dtype == MVLT_F32)
    launch_k(joint_embed_kernel<6, float, float>, dim3(grid), dim3(256), 0, stream, (const float*)feat, img_index, ids, text_mask, image_mask, word_emb, typepos, (float*)out, (bf16*)out_bf16_copy, kmask, B, n_obj, L, D, cls_id, sep_id, vocab_rows);
  else if (out_dtype == MVLT_BF16)
    launch_k(joint_embed_kernel<6, bf16, bf16>, dim3(grid), dim3(256), 0, stream, (const bf16*)feat, img_index, ids, text_mask, image_mask, word_emb, typepos, (bf16*)out, (bf16*)out_bf16_copy, kmask, B, n_obj, L, D, cls_id, sep_id, vocab_rows);
  else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_vit_embed(const float* patches, const float* cls, const float* pos, float* out, int B, int n_patch, int D,
                              cudaStream_t stream) {
  if (!patches || !cls || !pos || !out || B <= 0 || n_patch <= 0 || D <= 0 || D % 4) return MVLT_ERR_INVALID;
  if (((uintptr_t)patches & 15) || ((uintptr_t)cls & 15) || ((uintptr_t)pos & 15) || ((uintptr_t)out & 15)) return MVLT_ERR_INVALID;
  const long long total = (long long)B * (n_patch + 1) * (D / 4);
  long long g = (total + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  launch_k(vit_embed_kernel, dim3((unsigned)g), dim3(256), 0, stream, patches, cls, pos, out, B, n_patch, D);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
