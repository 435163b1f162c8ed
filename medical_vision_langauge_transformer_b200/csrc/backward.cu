// First slice of the training step (run_pretrain.py:177-184 `loss.backward()`): the row / attention kernels
// that the backward of ONE BertLayer (HF modeling_bert.py:359-421) needs around the tcgen05 GEMM, which computes every dgrad
// (dX = dY . W) and wgrad (dW = dY^T . X) as C = A . B^T on transposed bf16 copies:
//
//   mvlt_transpose_to_bf16      [rows, cols] fp32 | bf16 -> [cols, ld >= rows] bf16 (zero padded): the M-contiguous operands of wgrad
//   mvlt_layernorm_bwd_rows     dx, per-CTA partial dgamma / dbeta of LayerNorm(x) (torch.nn.functional.layer_norm backward)
//   mvlt_colsum                 bias gradients: column sums in a fixed order (deterministic two-stage reduction)
//   mvlt_gelu_bwd               du = df * (Phi(u) + u phi(u))  (erf GELU of HF:330-342)
//   mvlt_joint_attention_bwd    dq, dk, dv of softmax(q.k^T / 8 + mask) . v per (sample, head), probabilities recomputed
//
// These are plain CUDA-core kernels (fp32 math, bf16 storage): correct, deterministic, HBM- / FMA-bound — NOT yet the tensor-core
// versions the forward has; DESIGN.md §8 says what is and is not covered.  Parity: tests/test_backward_gpu.py against autograd of
// the oracle's bert_layer.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace mvlt {

// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
transpose_to_bf16_kernel(const T* __restrict__ in, long long ld_in, bf16* __restrict__ out, long long ld_out, long long rows, int cols) {
  __shared__ float tile[32][33];
  pdl_grid_sync();
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows of the tile per pass
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long r = r0 + ty + 8 * k;
    const int c = c0 + tx;
    tile[ty + 8 * k][tx] = (r < rows && c < cols) ? to_f32(in[r * ld_in + c]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const long long r = r0 + tx;
    if (c < cols && r < ld_out) out[(long long)c * ld_out + r] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row (C = 128 * NV, NV <= 8), rows grid-strided so that a CTA's column partials of
// dgamma = sum dy * xhat and dbeta = sum dy stay in registers until the end: part[blockIdx.x][0 | 1][C].
constexpr int LNB_MAX_NV = 8;
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma, float eps,
                     float* __restrict__ dx, bf16* __restrict__ dx_bf16, float* __restrict__ part, long long rows, int C) {
  __shared__ float red[8][128];
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int nv = NV;
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (v < nv) g[v] = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
    ag[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.0f / (float)C;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    float4 xv[NV], dv[NV];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (v < nv) {
        xv[v] = *(reinterpret_cast<const float4*>(x + row * C) + v * 32 + lane);
        dv[v] = *(reinterpret_cast<const float4*>(dy + row * C) + v * 32 + lane);
        s += (xv[v].x + xv[v].y) + (xv[v].z + xv[v].w);
      }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (v < nv) {
        xv[v].x -= mean; xv[v].y -= mean; xv[v].z -= mean; xv[v].w -= mean;
        q += xv[v].x * xv[v].x + xv[v].y * xv[v].y + xv[v].z * xv[v].z + xv[v].w * xv[v].w;
      }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * inv_c + eps);
    float c1 = 0.f, c2 = 0.f;     // sum of dxhat, sum of dxhat * xhat
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (v < nv) {
        xv[v].x *= rstd; xv[v].y *= rstd; xv[v].z *= rstd; xv[v].w *= rstd;                    // xhat
        ab[v].x += dv[v].x; ab[v].y += dv[v].y; ab[v].z += dv[v].z; ab[v].w += dv[v].w;
        ag[v].x = fmaf(dv[v].x, xv[v].x, ag[v].x); ag[v].y = fmaf(dv[v].y, xv[v].y, ag[v].y);
        ag[v].z = fmaf(dv[v].z, xv[v].z, ag[v].z); ag[v].w = fmaf(dv[v].w, xv[v].w, ag[v].w);
        dv[v].x *= g[v].x; dv[v].y *= g[v].y; dv[v].z *= g[v].z; dv[v].w *= g[v].w;            // dxhat
        c1 += (dv[v].x + dv[v].y) + (dv[v].z + dv[v].w);
        c2 += dv[v].x * xv[v].x + dv[v].y * xv[v].y + dv[v].z * xv[v].z + dv[v].w * xv[v].w;
      }
    c1 = warp_sum(c1) * inv_c;
    c2 = warp_sum(c2) * inv_c;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (v < nv) {
        float4 o;
        o.x = rstd * (dv[v].x - c1 - xv[v].x * c2); o.y = rstd * (dv[v].y - c1 - xv[v].y * c2);
        o.z = rstd * (dv[v].z - c1 - xv[v].z * c2); o.w = rstd * (dv[v].w - c1 - xv[v].w * c2);
        *(reinterpret_cast<float4*>(dx + row * C) + v * 32 + lane) = o;
        if (dx_bf16 != nullptr) store4(dx_bf16 + row * C + (v * 32 + lane) * 4, o);
      }
  }
  // the CTA's eight warps are added in warp order, 128 columns at a time
#pragma unroll
  for (int which = 0; which < 2; ++which)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 a = which == 0 ? ag[v] : ab[v];
      __syncthreads();
      red[warp][lane * 4] = a.x; red[warp][lane * 4 + 1] = a.y; red[warp][lane * 4 + 2] = a.z; red[warp][lane * 4 + 3] = a.w;
      __syncthreads();
      if (threadIdx.x < 128) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        part[((long long)blockIdx.x * 2 + which) * C + v * 128 + threadIdx.x] = t;
      }
    }
}

// out[j] = sum over p of part[p * stride + j] in index order (second stage of the deterministic column reductions)
__global__ void __launch_bounds__(256)
reduce_parts_kernel(const float* __restrict__ part, long long stride, int nparts, float* __restrict__ out, int n) {
  pdl_grid_sync();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  float t = 0.f;
  for (int p = 0; p < nparts; ++p) t += part[p * stride + j];
  out[j] = t;
}

// column sums of x[rows, cols]: slab s = rows [s * slab_rows, +slab_rows) -> part[s][cols]
template <typename T>
__global__ void __launch_bounds__(256)
colsum_slab_kernel(const T* __restrict__ x, long long ld, float* __restrict__ part, long long rows, int cols, int slab_rows) {
  pdl_grid_sync();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * slab_rows;
  const long long r1 = r0 + slab_rows < rows ? r0 + slab_rows : rows;
  float t = 0.f;
  for (long long r = r0; r < r1; ++r) t += to_f32(x[r * ld + c]);
  part[(long long)blockIdx.y * cols + c] = t;
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const bf16* __restrict__ u, const bf16* __restrict__ df, bf16* __restrict__ du, long long n4) {
  pdl_grid_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 uu = load4(u + i * 4), dd = load4(df + i * 4);
  auto d = [](float x) { return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x); };
  store4(du + i * 4, make_float4(dd.x * d(uu.x), dd.y * d(uu.y), dd.z * d(uu.z), dd.w * d(uu.w)));
}

// ------------------------------------------------------------------------------------------------------------------
// Attention backward for one (sample, head) per CTA, head_dim 64, S <= 160.  q, k, v, dO rows live in shared memory as bf16
// pairs with a 33-word row stride (a lane per key reads its own row conflict-free); phase A (a warp per query row) recomputes
// P, forms dP = dO . V^T, D = sum P * dP, dS = P * (dP - D), parks P and dS as bf16 [S][S] and writes dQ = dS . K / 8; phase B (a
// warp per key row) writes dV = P^T . dO and dK = dS^T . Q / 8.  Every sum runs in a fixed order.
constexpr int AB_MAX_S = 160;
constexpr int AB_THREADS = 256;
struct AttnBwdParams {
  const bf16* qkv;      // [B*S, 3C]
  const bf16* dctx;     // [B*S, C]
  const float* kmask;   // [B, S] additive (ignored when seq2seq)
  bf16* dqkv;           // [B*S, 3C]
  int S, C, seq2seq, obj_end;
  float scale;
};

__global__ void __launch_bounds__(AB_THREADS)
joint_attn_bwd_kernel(AttnBwdParams p) {
  extern __shared__ uint32_t sm[];
  pdl_grid_sync();
  const int S = p.S, head = blockIdx.x, b = blockIdx.y;
  const int ldp = (S + 2) | 1;                       // bf16 elements per P / dS row (odd word count is not needed: broadcast reads)
  uint32_t* Qs = sm;                                 // [S][33] words = 64 bf16 + pad
  uint32_t* Ks = Qs + S * 33;
  uint32_t* Vs = Ks + S * 33;
  uint32_t* Os = Vs + S * 33;
  bf16* Pm = reinterpret_cast<bf16*>(Os + S * 33);   // [S][ldp]
  bf16* Dm = Pm + (size_t)S * ldp;                   // [S][ldp]
  float* mk = reinterpret_cast<float*>(Pm + 2 * (size_t)S * ldp);   // [S] additive key mask (2 S ldp bf16 = a whole number of words)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)b * S;
  const int C3 = 3 * p.C;
  for (int idx = threadIdx.x; idx < S * 32; idx += AB_THREADS) {
    const int r = idx >> 5, w = idx & 31;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.qkv + (row0 + r) * C3 + head * 64) + w;
    Qs[r * 33 + w] = src[0];
    Ks[r * 33 + w] = src[p.C / 2];
    Vs[r * 33 + w] = src[p.C];
    Os[r * 33 + w] = reinterpret_cast<const uint32_t*>(p.dctx + (row0 + r) * p.C + head * 64)[w];
  }
  for (int j = threadIdx.x; j < S; j += AB_THREADS) mk[j] = p.seq2seq ? 0.f : p.kmask[(long long)b * S + j];
  __syncthreads();

  constexpr int JMAX = AB_MAX_S / 32;
  // ---- phase A: a warp per query row
  for (int i = warp; i < S; i += AB_THREADS / 32) {
    float sc[JMAX], dp[JMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < JMAX; ++t) {
      const int j = lane + 32 * t;
      sc[t] = -INFINITY; dp[t] = 0.f;
      if (j < S) {
        float a = 0.f, d = 0.f;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) {
          const float2 q = unpack_bf16x2(Qs[i * 33 + w]), k = unpack_bf16x2(Ks[j * 33 + w]);
          const float2 o = unpack_bf16x2(Os[i * 33 + w]), v = unpack_bf16x2(Vs[j * 33 + w]);
          a = fmaf(q.x, k.x, a); a = fmaf(q.y, k.y, a);
          d = fmaf(o.x, v.x, d); d = fmaf(o.y, v.y, d);
        }
        const float m = p.seq2seq ? ((j <= i || j <= p.obj_end) ? 0.f : -10000.f) : mk[j];
        sc[t] = a * p.scale + m;
        dp[t] = d;
        mx = fmaxf(mx, sc[t]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < JMAX; ++t) { sc[t] = lane + 32 * t < S ? __expf(sc[t] - mx) : 0.f; sum += sc[t]; }
    const float inv = 1.0f / warp_sum(sum);
    float dd = 0.f;
#pragma unroll
    for (int t = 0; t < JMAX; ++t) { sc[t] *= inv; dd = fmaf(sc[t], dp[t], dd); }
    dd = warp_sum(dd);
#pragma unroll
    for (int t = 0; t < JMAX; ++t) {
      const int j = lane + 32 * t;
      if (j < S) {
        Pm[(size_t)i * ldp + j] = __float2bfloat16_rn(sc[t]);
        Dm[(size_t)i * ldp + j] = __float2bfloat16_rn(sc[t] * (dp[t] - dd));
      }
    }
    __syncwarp();
    // dQ_i[d] = scale * sum_j dS_ij K_j[d]: the lane owns the bf16 pair `lane` of the row
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < S; ++j) {
      const float ds = __bfloat162float(Dm[(size_t)i * ldp + j]);
      const float2 k = unpack_bf16x2(Ks[j * 33 + lane]);
      acc.x = fmaf(ds, k.x, acc.x); acc.y = fmaf(ds, k.y, acc.y);
    }
    reinterpret_cast<uint32_t*>(p.dqkv + (row0 + i) * C3 + head * 64)[lane] = pack_bf16x2(acc.x * p.scale, acc.y * p.scale);
  }
  __syncthreads();
  // ---- phase B: a warp per key row
  for (int j = warp; j < S; j += AB_THREADS / 32) {
    float2 dk = make_float2(0.f, 0.f), dv = make_float2(0.f, 0.f);
    for (int i = 0; i < S; ++i) {
      const float pij = __bfloat162float(Pm[(size_t)i * ldp + j]), ds = __bfloat162float(Dm[(size_t)i * ldp + j]);
      const float2 q = unpack_bf16x2(Qs[i * 33 + lane]), o = unpack_bf16x2(Os[i * 33 + lane]);
      dk.x = fmaf(ds, q.x, dk.x); dk.y = fmaf(ds, q.y, dk.y);
      dv.x = fmaf(pij, o.x, dv.x); dv.y = fmaf(pij, o.y, dv.y);
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.dqkv + (row0 + j) * C3 + head * 64);
    dst[p.C / 2 + lane] = pack_bf16x2(dk.x * p.scale, dk.y * p.scale);
    dst[p.C + lane] = pack_bf16x2(dv.x, dv.y);
  }
}

static size_t attn_bwd_smem(int S) {
  const size_t ldp = (size_t)((S + 2) | 1);
  const size_t elems = 2 * (size_t)S * ldp;
  return 4 * (size_t)S * 33 * 4 + elems * 2 + (size_t)S * 4;
}

}  // namespace mvlt

using namespace mvlt;

extern "C" int mvlt_transpose_to_bf16(const void* in, int in_dtype, long long ld_in, void* out, long long ld_out, long long rows,
                                      int cols, cudaStream_t stream) {
  if (!in || !out || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < rows) return MVLT_ERR_INVALID;
  const dim3 grid((unsigned)((ld_out + 31) / 32), (unsigned)((cols + 31) / 32));
  if (in_dtype == MVLT_F32) launch_k(transpose_to_bf16_kernel<float>, grid, dim3(256), 0, stream, (const float*)in, ld_in, (bf16*)out, ld_out, rows, cols);
  else if (in_dtype == MVLT_BF16) launch_k(transpose_to_bf16_kernel<bf16>, grid, dim3(256), 0, stream, (const bf16*)in, ld_in, (bf16*)out, ld_out, rows, cols);
  else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

static int lnb_grid(long long rows) {
  const long long want = (rows + 7) / 8;
  return (int)(want < 296 ? want : 296);
}

// workspace of mvlt_layernorm_bwd_rows / mvlt_colsum: partial column sums, fp32
extern "C" long long mvlt_layernorm_bwd_workspace_bytes(long long rows, int C) { return (long long)lnb_grid(rows) * 2 * C * 4; }

// dx = d LayerNorm(x; gamma, beta, eps) / dx . dy per row (fp32, optional bf16 copy for the GEMMs that follow), dgamma = sum_rows
// dy * xhat, dbeta = sum_rows dy.  x is the PRE-normalisation input.  C % 128 == 0, C <= 1024; dense rows.
extern "C" int mvlt_layernorm_bwd_rows(const float* dy, const float* x, const float* gamma, float eps, float* dx, void* dx_bf16,
                                       float* dgamma, float* dbeta, void* workspace, long long rows, int C, cudaStream_t stream) {
  if (!dy || !x || !gamma || !dx || !dgamma || !dbeta || !workspace || rows <= 0) return MVLT_ERR_INVALID;
  if (C % 128 != 0 || C > 128 * LNB_MAX_NV) return MVLT_ERR_UNSUPPORTED;
  const int grid = lnb_grid(rows);
  float* part = reinterpret_cast<float*>(workspace);
#define MVLT_LNB_CASE(NV) case NV: launch_k(layernorm_bwd_kernel<NV>, dim3(grid), dim3(256), 0, stream, dy, x, gamma, eps, dx, (bf16*)dx_bf16, part, rows, C); break;
  switch (C / 128) {
    MVLT_LNB_CASE(1) MVLT_LNB_CASE(2) MVLT_LNB_CASE(3) MVLT_LNB_CASE(4) MVLT_LNB_CASE(5) MVLT_LNB_CASE(6) MVLT_LNB_CASE(7) MVLT_LNB_CASE(8)
    default: return MVLT_ERR_UNSUPPORTED;
  }
#undef MVLT_LNB_CASE
  MVLT_LAUNCH_CHECK();
  launch_k(reduce_parts_kernel, dim3((C + 255) / 256), dim3(256), 0, stream, (const float*)part, (long long)2 * C, grid, dgamma, C);
  MVLT_LAUNCH_CHECK();
  launch_k(reduce_parts_kernel, dim3((C + 255) / 256), dim3(256), 0, stream, (const float*)(part + C), (long long)2 * C, grid, dbeta, C);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

static int colsum_slabs(long long rows) {
  const long long s = (rows + 63) / 64;
  return (int)(s < 256 ? s : 256);
}
extern "C" long long mvlt_colsum_workspace_bytes(long long rows, int cols) { return (long long)colsum_slabs(rows) * cols * 4; }

// out[c] = sum_r x[r, c] (bias gradients), two stages in a fixed order
extern "C" int mvlt_colsum(const void* x, int dtype, long long ld, float* out, void* workspace, long long rows, int cols,
                           cudaStream_t stream) {
  if (!x || !out || !workspace || rows <= 0 || cols <= 0 || ld < cols) return MVLT_ERR_INVALID;
  const int slabs = colsum_slabs(rows);
  const int slab_rows = (int)((rows + slabs - 1) / slabs);
  float* part = reinterpret_cast<float*>(workspace);
  const dim3 grid((cols + 255) / 256, slabs);
  if (dtype == MVLT_F32) launch_k(colsum_slab_kernel<float>, grid, dim3(256), 0, stream, (const float*)x, ld, part, rows, cols, slab_rows);
  else if (dtype == MVLT_BF16) launch_k(colsum_slab_kernel<bf16>, grid, dim3(256), 0, stream, (const bf16*)x, ld, part, rows, cols, slab_rows);
  else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  launch_k(reduce_parts_kernel, dim3((cols + 255) / 256), dim3(256), 0, stream, (const float*)part, (long long)cols, slabs, out, cols);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

// du = df * gelu'(u), bf16 in / out, n % 4 == 0
extern "C" int mvlt_gelu_bwd(const void* u, const void* df, void* du, long long n, cudaStream_t stream) {
  if (!u || !df || !du || n <= 0 || n % 4 != 0) return MVLT_ERR_INVALID;
  launch_k(gelu_bwd_kernel, dim3((unsigned)((n / 4 + 255) / 256)), dim3(256), 0, stream, (const bf16*)u, (const bf16*)df, (bf16*)du, n / 4);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

// dqkv (bf16 [B*S, 3C]: dq | dk | dv) of ctx = softmax(q.k^T * scale + mask) . v given dctx (bf16 [B*S, C]); masks as in
// mvlt_joint_attention.  head_dim 64, S <= 160.
extern "C" int mvlt_joint_attention_bwd(const void* qkv, const float* kmask, const void* dctx, void* dqkv, int B, int S, int heads,
                                        int head_dim, int seq2seq, int obj_end, float scale, cudaStream_t stream) {
  if (!qkv || !dctx || !dqkv || B <= 0 || S <= 0 || heads <= 0 || (!seq2seq && !kmask)) return MVLT_ERR_INVALID;
  if (head_dim != 64 || S > AB_MAX_S) return MVLT_ERR_UNSUPPORTED;
  const size_t smem = attn_bwd_smem(S);
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(joint_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_bwd_smem(AB_MAX_S));
    if (e != cudaSuccess) return (int)e;
  }
  AttnBwdParams p;
  p.qkv = (const bf16*)qkv; p.dctx = (const bf16*)dctx; p.kmask = kmask; p.dqkv = (bf16*)dqkv;
  p.S = S; p.C = heads * 64; p.seq2seq = seq2seq; p.obj_end = obj_end; p.scale = scale;
  launch_k(joint_attn_bwd_kernel, dim3(heads, B), dim3(AB_THREADS), smem, stream, p);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
