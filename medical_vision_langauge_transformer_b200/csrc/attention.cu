// Fused small-sequence attention for both halves of the MVLT forward:
//
//   WINDOW mode  vfe.py:234-251 inside SwinTransformerBlock.forward :360-381 — per (image, 7x7 window, head):
//                S = scale*q.k^T + rel_pos_bias[h] (+ -100 shift mask), softmax, P.v.  The cyclic roll, window_partition,
//                window_reverse and the reverse roll (vfe.py:144-173, :361, :378) are NOT materialised: tokens stay in
//                natural [B, H*W, 3C] order and the CTA gathers/scatters rows through the index map
//                (b, wh, ww, i, j) -> (b, (wh*7+i+shift) mod H, (ww*7+j+shift) mod W).
//   JOINT mode   HF modeling_bert.py:115-140 (eager attention) with the MVLT masks of model.py:118-128,:162-183 — per
//                (sample, head): S = q.k^T/8 + additive key mask (0/-10000) or the seq2seq mask, softmax, P.v.
//
// bf16 path: one CTA per (group, head); Q/K/V rows staged in padded smem with 16 B loads, S and P.V on the tensor
// cores (mma.sync m16n8k16, fp32 accumulate) with the whole score row in registers — S/P never touch HBM.
// fp32 path ("parity mode"): same indexing, CUDA-core arithmetic.
#include "common.cuh"

namespace mvlt {

struct AttnParams {
  const void* qkv;   // [rows, ld_qkv]: q | k | v blocks of width C = heads*HD
  void* out;         // [rows, ld_out]
  long long ld_qkv, ld_out;
  int C, heads, ntok;
  float scale;
  // window mode
  int H, W, ws, shift;
  const float* relbias;  // [heads, 64, 64] fp32, zero padded
  // joint mode
  const float* kmask;    // [B, ntok]
  int seq2seq, obj_end;
};

template <bool WINDOW>
__device__ __forceinline__ long long token_row(const AttnParams& p, int group, int i) {
  if (WINDOW) {
    const int nww = p.W / p.ws, nw = (p.H / p.ws) * nww;
    const int b = group / nw, w = group % nw;
    const int h = ((w / nww) * p.ws + i / p.ws + p.shift) % p.H;
    const int x = ((w % nww) * p.ws + i % p.ws + p.shift) % p.W;
    return ((long long)b * p.H + h) * p.W + x;
  }
  return (long long)group * p.ntok + i;
}

// region id of token i of window w in the SHIFTED image (vfe.py:321-339)
__device__ __forceinline__ int shift_region(const AttnParams& p, int group, int i) {
  const int nww = p.W / p.ws, nw = (p.H / p.ws) * nww;
  const int w = group % nw;
  const int hs = (w / nww) * p.ws + i / p.ws, xs = (w % nww) * p.ws + i % p.ws;
  const int rh = hs < p.H - p.ws ? 0 : (hs < p.H - p.shift ? 1 : 2);
  const int rw = xs < p.W - p.ws ? 0 : (xs < p.W - p.shift ? 1 : 2);
  return rh * 3 + rw;
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// grid = (groups, heads); block = 32 * NPAD/16 (one 16-row query tile per warp)
template <int HD, int NPAD, bool WINDOW>
__global__ void __launch_bounds__(NPAD * 2)
attn_mma_kernel(const AttnParams p) {
  constexpr int LDS = HD + 8;  // padded row (bf16 elements): conflict-free ldmatrix
  constexpr int NT = NPAD / 8; // key tiles of 8
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_attn);
  bf16* Ks = Qs + NPAD * LDS;
  bf16* Vs = Ks + NPAD * LDS;
  float* aux = reinterpret_cast<float*>(Vs + NPAD * LDS);  // window: region ids, joint: key mask   [NPAD]

  const int group = blockIdx.x, head = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bf16* qkv = reinterpret_cast<const bf16*>(p.qkv);

  // ---- stage Q, K, V rows (16 B chunks), zero-fill the padding rows ----
  constexpr int CPR = HD / 8;  // 16 B chunks per row per matrix
  for (int idx = tid; idx < NPAD * 3 * CPR; idx += blockDim.x) {
    const int ch = idx % CPR, which = (idx / CPR) % 3, i = idx / (3 * CPR);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (i < p.ntok) {
      const long long row = token_row<WINDOW>(p, group, i);
      val = *reinterpret_cast<const uint4*>(qkv + row * p.ld_qkv + which * p.C + head * HD + ch * 8);
    }
    bf16* dst = (which == 0 ? Qs : which == 1 ? Ks : Vs) + i * LDS + ch * 8;
    *reinterpret_cast<uint4*>(dst) = val;
  }
  for (int i = tid; i < NPAD; i += blockDim.x) {
    float a = 0.f;
    if (i < p.ntok) {
      if (WINDOW) a = p.shift > 0 ? (float)shift_region(p, group, i) : 0.f;
      else a = p.kmask[(long long)group * p.ntok + i];
    }
    aux[i] = a;
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16;  // first query row of this warp
  if (r0 < p.ntok) {
    // ---- S = Q K^T ----
    uint32_t qa[HD / 16][4];
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      ldsm_x4(smem_u32(Qs + (r0 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kq = 0; kq < HD / 32; ++kq) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(Ks + (nt * 8 + (lane & 7)) * LDS + kq * 32 + (lane >> 3) * 8), b0, b1, b2, b3);
        mma_bf16_16816(s[nt], qa[2 * kq][0], qa[2 * kq][1], qa[2 * kq][2], qa[2 * kq][3], b0, b1);
        mma_bf16_16816(s[nt], qa[2 * kq + 1][0], qa[2 * kq + 1][1], qa[2 * kq + 1][2], qa[2 * kq + 1][3], b2, b3);
      }
    }
    // ---- scale, bias, mask; row softmax (rows r0+g and r0+g+8) ----
    const int i0 = r0 + g, i1 = r0 + g + 8;
    float mx0 = -INFINITY, mx1 = -INFINITY;
    float ra0 = 0.f, ra1 = 0.f;
    const float* rb0 = nullptr;
    const float* rb1 = nullptr;
    if (WINDOW) {
      ra0 = aux[i0]; ra1 = aux[i1];
      rb0 = p.relbias + ((long long)head * 64 + i0) * 64;
      rb1 = p.relbias + ((long long)head * 64 + i1) * 64;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int j = nt * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = j + e;
        float add0, add1;
        if (WINDOW) {
          const float rj = aux[jj];
          add0 = __ldg(rb0 + jj) + (rj != ra0 ? -100.f : 0.f);
          add1 = __ldg(rb1 + jj) + (rj != ra1 ? -100.f : 0.f);
        } else if (p.seq2seq) {
          add0 = (jj <= i0 || jj <= p.obj_end) ? 0.f : -10000.f;
          add1 = (jj <= i1 || jj <= p.obj_end) ? 0.f : -10000.f;
        } else {
          add0 = add1 = aux[jj];
        }
        const bool valid = jj < p.ntok;
        s[nt][e] = valid ? s[nt][e] * p.scale + add0 : -INFINITY;
        s[nt][2 + e] = valid ? s[nt][2 + e] * p.scale + add1 : -INFINITY;
        mx0 = fmaxf(mx0, s[nt][e]);
        mx1 = fmaxf(mx1, s[nt][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    constexpr float LOG2E = 1.4426950408889634f;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = exp2f((s[nt][e] - mx0) * LOG2E);
        s[nt][2 + e] = exp2f((s[nt][2 + e] - mx1) * LOG2E);
        sum0 += s[nt][e];
        sum1 += s[nt][2 + e];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // ---- O = P V ----
    float o[HD / 8][4];
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NPAD / 16; ++kk) {
      const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      const uint32_t a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      const uint32_t a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * dp], a0, a1, a2, a3, b0, b1);
        mma_bf16_16816(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // ---- normalise, park in this warp's own Q rows, then 16 B row-contiguous stores ----
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    __syncwarp();
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      *reinterpret_cast<uint32_t*>(Qs + (r0 + g) * LDS + dn * 8 + 2 * t) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      *reinterpret_cast<uint32_t*>(Qs + (r0 + g + 8) * LDS + dn * 8 + 2 * t) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
    __syncwarp();
    bf16* out = reinterpret_cast<bf16*>(p.out);
    for (int idx = lane; idx < 16 * CPR; idx += 32) {
      const int rr = idx / CPR, ch = idx % CPR;
      const int i = r0 + rr;
      if (i < p.ntok) {
        const long long row = token_row<WINDOW>(p, group, i);
        *reinterpret_cast<uint4*>(out + row * p.ld_out + head * HD + ch * 8) =
            *reinterpret_cast<const uint4*>(Qs + i * LDS + ch * 8);
      }
    }
  }
}

// ---- fp32 parity path: CUDA cores, scores in smem -------------------------------------------------------------------
template <int HD, bool WINDOW>
__global__ void __launch_bounds__(256)
attn_f32_kernel(const AttnParams p, int npad) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int N = p.ntok;
  constexpr int LD = HD + 1;
  float* Qs = reinterpret_cast<float*>(smem_attn);
  float* Ks = Qs + npad * LD;
  float* Vs = Ks + npad * LD;
  float* aux = Vs + npad * LD;
  float* S = aux + npad;  // [N][npad+1]
  const int lds = npad + 1;
  const int group = blockIdx.x, head = blockIdx.y, tid = threadIdx.x;
  const float* qkv = reinterpret_cast<const float*>(p.qkv);
  for (int idx = tid; idx < N * 3 * HD; idx += blockDim.x) {
    const int d = idx % HD, which = (idx / HD) % 3, i = idx / (3 * HD);
    const long long row = token_row<WINDOW>(p, group, i);
    (which == 0 ? Qs : which == 1 ? Ks : Vs)[i * LD + d] = qkv[row * p.ld_qkv + which * p.C + head * HD + d];
  }
  for (int i = tid; i < N; i += blockDim.x)
    aux[i] = WINDOW ? (p.shift > 0 ? (float)shift_region(p, group, i) : 0.f) : p.kmask[(long long)group * N + i];
  __syncthreads();
  for (int idx = tid; idx < N * N; idx += blockDim.x) {
    const int i = idx / N, j = idx % N;
    float acc = 0.f;
    if (WINDOW) {
      // reference scales q before the product (vfe.py:234)
#pragma unroll
      for (int d = 0; d < HD; ++d) acc = fmaf(Qs[i * LD + d] * p.scale, Ks[j * LD + d], acc);
      acc += p.relbias[((long long)head * 64 + i) * 64 + j];
      if (aux[i] != aux[j]) acc += -100.f;
    } else {
#pragma unroll
      for (int d = 0; d < HD; ++d) acc = fmaf(Qs[i * LD + d], Ks[j * LD + d], acc);
      acc *= p.scale;
      acc += p.seq2seq ? ((j <= i || j <= p.obj_end) ? 0.f : -10000.f) : aux[j];
    }
    S[i * lds + j] = acc;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int i = warp; i < N; i += nwarps) {
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, S[i * lds + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = expf(S[i * lds + j] - mx);
      S[i * lds + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) S[i * lds + j] *= inv;
  }
  __syncthreads();
  float* out = reinterpret_cast<float*>(p.out);
  for (int idx = tid; idx < N * HD; idx += blockDim.x) {
    const int i = idx / HD, d = idx % HD;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(S[i * lds + j], Vs[j * LD + d], acc);
    out[token_row<WINDOW>(p, group, i) * p.ld_out + head * HD + d] = acc;
  }
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return MVLT_OK;
}

constexpr int JOINT_NPAD = 144;   // 131 joint tokens (L=80) -> 9 query tiles; shorter L uses the same kernel
constexpr int JOINT_NPAD_S = 96;  // S <= 96 (e.g. SLAKE L=23 -> 74, VQA-RAD L=30 -> 81)

}  // namespace mvlt

using namespace mvlt;

extern "C" int mvlt_attn_init(void) {
  int rc;
  if ((rc = set_smem(attn_mma_kernel<64, JOINT_NPAD, false>, 3 * JOINT_NPAD * 72 * 2 + JOINT_NPAD * 4)) != MVLT_OK) return rc;
  if ((rc = set_smem(attn_f32_kernel<64, false>, 200 * 1024)) != MVLT_OK) return rc;
  if ((rc = set_smem(attn_f32_kernel<32, true>, 64 * 1024)) != MVLT_OK) return rc;
  return MVLT_OK;
}

// Swin window attention on tokens kept in natural order.  qkv: [B*H*W, 3C], out: [B*H*W, C].
extern "C" int mvlt_window_attention(const void* qkv, void* out, int dtype, const float* relbias, int B, int H, int W,
                                     int C, int heads, int window, int shift, float scale, cudaStream_t stream) {
  if (!qkv || !out || !relbias || B <= 0 || heads <= 0 || C != heads * 32) return MVLT_ERR_INVALID;
  if (window != 7 || H % window || W % window || shift < 0 || shift >= window) return MVLT_ERR_UNSUPPORTED;
  AttnParams p{};
  p.qkv = qkv; p.out = out; p.ld_qkv = 3LL * C; p.ld_out = C; p.C = C; p.heads = heads; p.ntok = window * window;
  p.scale = scale; p.H = H; p.W = W; p.ws = window; p.shift = shift; p.relbias = relbias;
  dim3 grid(B * (H / window) * (W / window), heads);
  if (dtype == MVLT_BF16) {
    constexpr int NPAD = 64;
    attn_mma_kernel<32, NPAD, true><<<grid, NPAD * 2, 3 * NPAD * 40 * 2 + NPAD * 4, stream>>>(p);
  } else if (dtype == MVLT_F32) {
    const int npad = 52;
    const int bytes = (3 * npad * 33 + npad + 49 * (npad + 1)) * 4;
    attn_f32_kernel<32, true><<<grid, 128, bytes, stream>>>(p, npad);
  } else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

// BERT joint attention.  qkv: [B*S, 3*768] (q|k|v), out: [B*S, 768]; kmask: [B,S] additive (ignored when seq2seq).
extern "C" int mvlt_joint_attention(const void* qkv, void* out, int dtype, const float* kmask, int B, int S, int heads,
                                    int head_dim, int seq2seq, int obj_end, float scale, cudaStream_t stream) {
  if (!qkv || !out || !kmask || B <= 0 || S <= 0 || heads <= 0) return MVLT_ERR_INVALID;
  if (head_dim != 64) return MVLT_ERR_UNSUPPORTED;
  const int C = heads * head_dim;
  AttnParams p{};
  p.qkv = qkv; p.out = out; p.ld_qkv = 3LL * C; p.ld_out = C; p.C = C; p.heads = heads; p.ntok = S; p.scale = scale;
  p.kmask = kmask; p.seq2seq = seq2seq; p.obj_end = obj_end;
  dim3 grid(B, heads);
  if (dtype == MVLT_BF16) {
    if (S <= JOINT_NPAD_S)
      attn_mma_kernel<64, JOINT_NPAD_S, false><<<grid, JOINT_NPAD_S * 2, 3 * JOINT_NPAD_S * 72 * 2 + JOINT_NPAD_S * 4, stream>>>(p);
    else if (S <= JOINT_NPAD)
      attn_mma_kernel<64, JOINT_NPAD, false><<<grid, JOINT_NPAD * 2, 3 * JOINT_NPAD * 72 * 2 + JOINT_NPAD * 4, stream>>>(p);
    else return MVLT_ERR_UNSUPPORTED;
  } else if (dtype == MVLT_F32) {
    const int npad = (S + 3) & ~3;
    const int bytes = (3 * npad * 65 + npad + S * (npad + 1)) * 4;
    if (bytes > 200 * 1024) return MVLT_ERR_UNSUPPORTED;
    attn_f32_kernel<64, false><<<grid, 256, bytes, stream>>>(p, npad);
  } else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
