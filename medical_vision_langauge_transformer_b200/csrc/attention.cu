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
#include <stdlib.h>

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
  int head_major;        // joint mode: blockIdx.x = head (the 12 heads of a sample run together: whole qkv rows stay hot)
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E = 1.4426950408889634f;
constexpr float NEG_BIG = -1.0e30f;  // "minus infinity" that stays finite through (x - max) and scale multiplies

// ---- Swin window attention, bf16: ONE WARP per (image, window, head) ------------------------------------------------
// The 49x32 q, k, v slices of the item are cp.async'ed into this warp's private 12 KB of shared memory (64-byte rows,
// 16-byte chunks XOR-swizzled by (row >> 1) & 3: conflict-free ldmatrix without padding -> 16 warps per SM), rows found
// through the roll/partition index map computed once per item.  K fragments stay in registers across the four 16-row
// query tiles.  The score accumulators are INITIALISED from a host-built table that already holds
// (rel-pos bias + shift mask) / scale in the mma C-fragment layout, one float4 per (class, head, m-tile, n-tile, lane),
// with -1e30 in the columns of the non-existent keys 49..55: no bias gather, no mask logic, no scaling in the kernel,
// and exp2(scale*log2e * acc - max) is one FFMA + one MUFU per score.  P.V runs from the S fragments in registers.
// No __syncthreads anywhere: the warps of a CTA are independent pipelines (loads of one overlap math of another).
constexpr int WA_ROW = 32;                                   // bf16 elements per staged row (64 B)
constexpr int WA_MAT = 64 * WA_ROW;                          // elements per staged matrix (4 KB)
constexpr int WA_WARP_BYTES = 3 * WA_MAT * 2 + 64 * 4;       // Q, K, V + row index table
constexpr int WA_WARPS = 8;
constexpr int WA_CTAS_PER_SM = 2;

// element offset of 16-byte chunk `chunk` (0..3) of row `row` in a staged matrix
__device__ __forceinline__ int wa_off(int row, int chunk) { return row * WA_ROW + ((chunk ^ ((row >> 1) & 3)) << 3); }

struct WinParams {
  const bf16* qkv;
  bf16* out;
  const float4* bias_frag;  // [n_cls, heads, 4 m-tiles, 7 n-tiles, 32 lanes] float4 (see ops.window_bias_fragments)
  int H, W, C, heads, shift, n_items, nWh, nWw;
  float scale;
};

__global__ void __launch_bounds__(WA_WARPS * 32, WA_CTAS_PER_SM)
window_attn_warp_kernel(const WinParams p) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  bf16* Qs = reinterpret_cast<bf16*>(smem_attn + warp * WA_WARP_BYTES);
  bf16* Ks = Qs + WA_MAT;
  bf16* Vs = Ks + WA_MAT;
  int* rowoff = reinterpret_cast<int*>(Vs + WA_MAT);   // token i of this window -> row of the [B*H*W, .] matrices
  // padding rows (tokens 49..63) are zero for the whole kernel: loads only ever touch rows 0..48
  for (int i = lane; i < 3 * WA_MAT / 8; i += 32) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0, 0, 0, 0);
  for (int i = lane; i < 64; i += 32) rowoff[i] = 0;
  __syncwarp();
  pdl_grid_sync();  // the shared-memory setup above overlapped the previous kernel's tail

  const int nW = p.nWh * p.nWw;
  const long long ld_qkv = 3LL * p.C, ld_out = p.C;
  const float c = p.scale * LOG2E;
  for (int item = blockIdx.x * WA_WARPS + warp; item < p.n_items; item += gridDim.x * WA_WARPS) {
    const int head = item % p.heads, wg = item / p.heads;
    const int b = wg / nW, w = wg - b * nW;
    const int wh = w / p.nWw, ww = w - wh * p.nWw;
    // window class of the shift mask (vfe.py:321-339): windows of the last window row / column straddle the roll seam
    const int cls = p.shift > 0 ? ((wh == p.nWh - 1 ? 2 : 0) | (ww == p.nWw - 1 ? 1 : 0)) : 0;
    for (int i = lane; i < 49; i += 32) {
      const int r = i / 7, cc = i - r * 7;
      int h = wh * 7 + r + p.shift, x = ww * 7 + cc + p.shift;   // torch.roll(-shift) then partition == read at +shift
      if (h >= p.H) h -= p.H;
      if (x >= p.W) x -= p.W;
      rowoff[i] = (b * p.H + h) * p.W + x;
    }
    __syncwarp();
    {  // 49 rows x (4 q + 4 k + 4 v) 16-byte chunks: 16 lanes per row, 2 rows per pass
      const int ch = lane & 15, which = ch >> 2, c4 = ch & 3;
      const bf16* src0 = p.qkv + (long long)which * p.C + head * 32 + c4 * 8;
      bf16* dst0 = Qs + which * WA_MAT;
      if (ch < 12) {
#pragma unroll 5
        for (int it = 0; it < 25; ++it) {
          const int row = 2 * it + (lane >> 4);
          if (row < 49) cp_async16(dst0 + wa_off(row, c4), src0 + (long long)rowoff[row] * ld_qkv);
        }
      }
      cp_async_wait_all();
    }
    __syncwarp();

    uint32_t kf[7][4];  // B fragments of K^T for key tiles 0..6 (keys 0..55), both k-steps of the 32-wide head
#pragma unroll
    for (int nt = 0; nt < 7; ++nt)
      ldsm_x4(smem_u32(Ks + wa_off(nt * 8 + (lane & 7), lane >> 3)), kf[nt][0], kf[nt][1], kf[nt][2], kf[nt][3]);
    const float4* bt = p.bias_frag + (long long)(cls * p.heads + head) * (4 * 7 * 32) + lane;

#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      const int r0 = mt * 16, i0 = r0 + g, i1 = r0 + g + 8;
      uint32_t qa[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ldsm_x4(smem_u32(Qs + wa_off(r0 + (lane & 15), ks * 2 + (lane >> 4))), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
      // The table values are fetched first (L2 latency: the kernel's largest single stall when the accumulators were
      // initialised from them), the MMAs run on zeroed accumulators meanwhile and the table is added afterwards; the two
      // k-steps are issued as two passes over the seven key tiles so that consecutive HMMAs never share an accumulator.
      float4 b4[7];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) b4[nt] = __ldg(bt + (mt * 7 + nt) * 32);
      float s[7][4];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) mma_bf16_16816(s[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kf[nt][0], kf[nt][1]);
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) mma_bf16_16816(s[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kf[nt][2], kf[nt][3]);
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) { s[nt][0] += b4[nt].x; s[nt][1] += b4[nt].y; s[nt][2] += b4[nt].z; s[nt][3] += b4[nt].w; }
      float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float m0c = -mx0 * c, m1c = -mx1 * c;
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        s[nt][0] = ex2_approx(fmaf(s[nt][0], c, m0c)); s[nt][1] = ex2_approx(fmaf(s[nt][1], c, m0c));
        s[nt][2] = ex2_approx(fmaf(s[nt][2], c, m1c)); s[nt][3] = ex2_approx(fmaf(s[nt][3], c, m1c));
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

      float o[4][4];
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        const uint32_t a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        const uint32_t a2 = kk < 3 ? pack_bf16x2(s[kk < 3 ? 2 * kk + 1 : 0][0], s[kk < 3 ? 2 * kk + 1 : 0][1]) : 0u;  // keys 56..63: P = 0
        const uint32_t a3 = kk < 3 ? pack_bf16x2(s[kk < 3 ? 2 * kk + 1 : 0][2], s[kk < 3 ? 2 * kk + 1 : 0][3]) : 0u;
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(Vs + wa_off(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4))), b0, b1, b2, b3);
          mma_bf16_16816(o[2 * dp], a0, a1, a2, a3, b0, b1);
          mma_bf16_16816(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
        }
      }
      // normalise and park the tile in its own Q rows (already in registers); rows >= 49 stay zero
      const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
      __syncwarp();
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        if (i0 < 49) *reinterpret_cast<uint32_t*>(Qs + wa_off(i0, dn) + 2 * t) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
        if (i1 < 49) *reinterpret_cast<uint32_t*>(Qs + wa_off(i1, dn) + 2 * t) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
      }
    }
    __syncwarp();
    {  // window_reverse + roll back == scatter through the same index map: 49 rows x 4 chunks of 16 B
      bf16* dst0 = p.out + head * 32 + (lane & 3) * 8;
#pragma unroll
      for (int it = 0; it < 7; ++it) {
        const int row = it * 8 + (lane >> 2);
        if (row < 49)
          *reinterpret_cast<uint4*>(dst0 + (long long)rowoff[row] * ld_out) = *reinterpret_cast<const uint4*>(Qs + wa_off(row, lane & 3));
      }
    }
    __syncwarp();  // the next item rewrites rowoff / Q / K / V
  }
}

// ---- BERT joint attention, bf16: one CTA per (sample, head), key-chunked online softmax ----------------------------------
// grid = (heads, B): the twelve heads of a sample are scheduled together.  Three warps per CTA, each walking the 16-row query
// tiles warp, warp + 3, ...; the keys are walked in chunks of 48 with a running (max, sum) per row — 24 score registers
// instead of the 72 of a full 144-key row; K / V live in UNPADDED, XOR-swizzled 128-byte rows, each query tile is staged in a
// per-warp 2 KB buffer that is reused for the output: 43.6 KB and 128 registers, four to five CTAs per SM.  The score
// accumulators start from (additive key mask | seq2seq mask) / scale; exp2 with the scale folded in; P.V from registers.
// History (ncu, profiles/r01_ncu_joint_attention_variants.txt): one warp per tile with full rows (9 warps allocated as 12,
// 96-register cap, 100 B of spills, ONE CTA per SM), then 3 warps x 3 tiles with full rows (168 registers, 3 CTAs), then this
// one all run 29-31 us for 768 (sample, head) items: every CTA of the single wave loads, then computes, in lock-step, so
// the kernel is the SUM of its load phase and its mma.sync phase whatever the occupancy; a persistent, double-buffered item
// loop is what would overlap them.
// (launch bound of 128 threads although 96 run: warps are allocated in fours, so the register cap must be computed for four)
template <int NCH>
__global__ void __launch_bounds__(128, NCH <= 3 ? 4 : (NCH == 4 ? 3 : 2))
joint_attn_flash_kernel(const AttnParams p) {
  pdl_grid_sync();
  constexpr int HD = 64, KEYS = 48 * NCH, NW = 3, CPR = HD / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_attn);
  bf16* Vs = Ks + KEYS * HD;
  bf16* Qall = Vs + KEYS * HD;
  float* aux = reinterpret_cast<float*>(Qall + NW * 16 * HD);  // [KEYS] additive key mask / scale (NEG_BIG beyond ntok)
  const int group = p.head_major ? blockIdx.y : blockIdx.x, head = p.head_major ? blockIdx.x : blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  bf16* Qw = Qall + warp * 16 * HD;
  const bf16* qkv = reinterpret_cast<const bf16*>(p.qkv) + (long long)group * p.ntok * p.ld_qkv + head * HD;
  const float inv_scale = 1.0f / p.scale, c = p.scale * LOG2E;
  // element offset of 16-byte chunk `ch` of row `row` in a swizzled [rows][64] bf16 tile
  auto swz = [](int row, int ch) { return row * HD + ((ch ^ (row & 7)) << 3); };

  for (int idx = tid; idx < KEYS * 2 * CPR; idx += NW * 32) {
    const int ch = idx % CPR, which = (idx / CPR) & 1, i = idx / (2 * CPR);
    bf16* dst = (which ? Vs : Ks) + swz(i, ch);
    if (i < p.ntok) cp_async16(dst, qkv + (long long)i * p.ld_qkv + (1 + which) * p.C + ch * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < KEYS; i += NW * 32)
    aux[i] = i < p.ntok ? (p.seq2seq ? 0.f : p.kmask[(long long)group * p.ntok + i] * inv_scale) : NEG_BIG;
  cp_async_wait_all();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  bf16* out = reinterpret_cast<bf16*>(p.out) + (long long)group * p.ntok * p.ld_out + head * HD;
  for (int r0 = warp * 16; r0 < p.ntok; r0 += NW * 16) {
    // ---- this warp's query tile -> Qw
    for (int idx = lane; idx < 16 * CPR; idx += 32) {
      const int rr = idx / CPR, ch = idx % CPR;
      bf16* dst = Qw + swz(rr, ch);
      if (r0 + rr < p.ntok) cp_async16(dst, qkv + (long long)(r0 + rr) * p.ld_qkv + ch * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    cp_async_wait_all();
    __syncwarp();
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldsm_x4(smem_u32(Qw + swz(lane & 15, ks * 2 + (lane >> 4))), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    const int i0 = r0 + g, i1 = r0 + g + 8;
    float o[HD / 8][4];
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    float m0 = NEG_BIG, m1 = NEG_BIG, l0 = 0.f, l1 = 0.f;   // running row max (score units) and this thread's partial row sums
#pragma unroll
    for (int chn = 0; chn < NCH; ++chn) {
      const int k0 = chn * 48;
      if (k0 < p.ntok) {   // warp-uniform; a chunk entirely beyond the sequence contributes nothing
        float s[6][4];
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const int j = k0 + nt * 8 + 2 * t;
          const float2 m = *reinterpret_cast<const float2*>(aux + j);
          s[nt][0] = s[nt][2] = m.x;
          s[nt][1] = s[nt][3] = m.y;
          if (p.seq2seq) {  // model.py:118-123: text rows see the image block and the text up to themselves
            const float blocked = -10000.f * inv_scale;
            if (j > i0 && j > p.obj_end) s[nt][0] += blocked;
            if (j + 1 > i0 && j + 1 > p.obj_end) s[nt][1] += blocked;
            if (j > i1 && j > p.obj_end) s[nt][2] += blocked;
            if (j + 1 > i1 && j + 1 > p.obj_end) s[nt][3] += blocked;
          }
        }
#pragma unroll
        for (int kq = 0; kq < 2; ++kq)
#pragma unroll
          for (int nt = 0; nt < 6; ++nt) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(smem_u32(Ks + swz(k0 + nt * 8 + (lane & 7), kq * 4 + (lane >> 3))), b0, b1, b2, b3);
            mma_bf16_16816(s[nt], qa[2 * kq][0], qa[2 * kq][1], qa[2 * kq][2], qa[2 * kq][3], b0, b1);
            mma_bf16_16816(s[nt], qa[2 * kq + 1][0], qa[2 * kq + 1][1], qa[2 * kq + 1][2], qa[2 * kq + 1][3], b2, b3);
          }
        float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
        const float corr0 = ex2_approx((m0 - n0) * c), corr1 = ex2_approx((m1 - n1) * c);
        m0 = n0; m1 = n1;
        const float m0c = -n0 * c, m1c = -n1 * c;
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pk[6][2];  // P as bf16 pairs: [nt][0] = row g, [nt][1] = row g+8
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const float p0 = ex2_approx(fmaf(s[nt][0], c, m0c)), p1 = ex2_approx(fmaf(s[nt][1], c, m0c));
          const float p2 = ex2_approx(fmaf(s[nt][2], c, m1c)), p3 = ex2_approx(fmaf(s[nt][3], c, m1c));
          sum0 += p0 + p1;
          sum1 += p2 + p3;
          pk[nt][0] = pack_bf16x2(p0, p1);
          pk[nt][1] = pack_bf16x2(p2, p3);
        }
        l0 = fmaf(l0, corr0, sum0);
        l1 = fmaf(l1, corr1, sum1);
        if (chn > 0) {
#pragma unroll
          for (int dn = 0; dn < HD / 8; ++dn) { o[dn][0] *= corr0; o[dn][1] *= corr0; o[dn][2] *= corr1; o[dn][3] *= corr1; }
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
#pragma unroll
          for (int dp = 0; dp < HD / 16; ++dp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(smem_u32(Vs + swz(k0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4))), b0, b1, b2, b3);
            mma_bf16_16816(o[2 * dp], pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1], b0, b1);
            mma_bf16_16816(o[2 * dp + 1], pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1], b2, b3);
          }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    __syncwarp();   // every lane has read its Q fragments: Qw becomes the output staging tile
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn) {
      *reinterpret_cast<uint32_t*>(Qw + swz(g, dn) + 2 * t) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      *reinterpret_cast<uint32_t*>(Qw + swz(g + 8, dn) + 2 * t) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
    __syncwarp();
    for (int idx = lane; idx < 16 * CPR; idx += 32) {
      const int rr = idx / CPR, ch = idx % CPR;
      if (r0 + rr < p.ntok)
        *reinterpret_cast<uint4*>(out + (long long)(r0 + rr) * p.ld_out + ch * 8) = *reinterpret_cast<const uint4*>(Qw + swz(rr, ch));
    }
    __syncwarp();   // the staging tile is free for the next query tile
  }
}

// ---- fp32 parity path: CUDA cores, scores in smem -------------------------------------------------------------------
// Query rows are processed in blocks of `rb` (gridDim.z blocks per (group, head)): K / V of the whole sequence plus the
// scores of one block live in shared memory, so joint sequences up to ~300 tokens (ViT / linear-patch backbones) fit.
template <int HD, bool WINDOW>
__global__ void __launch_bounds__(256)
attn_f32_kernel(const AttnParams p, int npad, int rb) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int N = p.ntok;
  constexpr int LD = HD + 1;
  float* Qs = reinterpret_cast<float*>(smem_attn);   // [rb][LD]
  float* Ks = Qs + rb * LD;                           // [npad][LD]
  float* Vs = Ks + npad * LD;
  float* aux = Vs + npad * LD;
  float* S = aux + npad;  // [rb][npad+1]
  const int lds = npad + 1;
  const int group = blockIdx.x, head = blockIdx.y, tid = threadIdx.x;
  const int row0 = blockIdx.z * rb, nr = min(rb, N - row0);
  const float* qkv = reinterpret_cast<const float*>(p.qkv);
  for (int idx = tid; idx < N * 2 * HD; idx += blockDim.x) {
    const int d = idx % HD, which = (idx / HD) % 2, i = idx / (2 * HD);
    const long long row = token_row<WINDOW>(p, group, i);
    (which == 0 ? Ks : Vs)[i * LD + d] = qkv[row * p.ld_qkv + (1 + which) * p.C + head * HD + d];
  }
  for (int idx = tid; idx < nr * HD; idx += blockDim.x) {
    const int d = idx % HD, i = idx / HD;
    Qs[i * LD + d] = qkv[token_row<WINDOW>(p, group, row0 + i) * p.ld_qkv + head * HD + d];
  }
  for (int i = tid; i < N; i += blockDim.x)
    aux[i] = WINDOW ? (p.shift > 0 ? (float)shift_region(p, group, i) : 0.f) : p.kmask[(long long)group * N + i];
  __syncthreads();
  for (int idx = tid; idx < nr * N; idx += blockDim.x) {
    const int il = idx / N, j = idx % N, i = row0 + il;
    float acc = 0.f;
    if (WINDOW) {
      // reference scales q before the product (vfe.py:234)
#pragma unroll
      for (int d = 0; d < HD; ++d) acc = fmaf(Qs[il * LD + d] * p.scale, Ks[j * LD + d], acc);
      acc += p.relbias[((long long)head * 64 + i) * 64 + j];
      if (aux[i] != aux[j]) acc += -100.f;
    } else {
#pragma unroll
      for (int d = 0; d < HD; ++d) acc = fmaf(Qs[il * LD + d], Ks[j * LD + d], acc);
      acc *= p.scale;
      acc += p.seq2seq ? ((j <= i || j <= p.obj_end) ? 0.f : -10000.f) : aux[j];
    }
    S[il * lds + j] = acc;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int i = warp; i < nr; i += nwarps) {
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
  for (int idx = tid; idx < nr * HD; idx += blockDim.x) {
    const int i = idx / HD, d = idx % HD;
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(S[i * lds + j], Vs[j * LD + d], acc);
    out[token_row<WINDOW>(p, group, row0 + i) * p.ld_out + head * HD + d] = acc;
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


}  // namespace mvlt

using namespace mvlt;

template <int NCH> constexpr int joint_flash_smem() { return 48 * NCH * 64 * 2 * 2 + 3 * 16 * 64 * 2 + 48 * NCH * 4; }

extern "C" int mvlt_attn_init(void) {
  static unsigned long long devices = 0;
  if (!first_use_on_device(devices)) return MVLT_OK;
  int rc;
  if ((rc = set_smem(joint_attn_flash_kernel<2>, joint_flash_smem<2>())) != MVLT_OK) return rc;
  if ((rc = set_smem(joint_attn_flash_kernel<3>, joint_flash_smem<3>())) != MVLT_OK) return rc;
  if ((rc = set_smem(joint_attn_flash_kernel<4>, joint_flash_smem<4>())) != MVLT_OK) return rc;
  if ((rc = set_smem(joint_attn_flash_kernel<5>, joint_flash_smem<5>())) != MVLT_OK) return rc;
  if ((rc = set_smem(joint_attn_flash_kernel<6>, joint_flash_smem<6>())) != MVLT_OK) return rc;
  if ((rc = set_smem(window_attn_warp_kernel, WA_WARPS * WA_WARP_BYTES)) != MVLT_OK) return rc;
  if ((rc = set_smem(attn_f32_kernel<64, false>, 200 * 1024)) != MVLT_OK) return rc;
  if ((rc = set_smem(attn_f32_kernel<32, true>, 64 * 1024)) != MVLT_OK) return rc;
  return MVLT_OK;
}

// Swin window attention on tokens kept in natural order.  qkv: [B*H*W, 3C], out: [B*H*W, C].
extern "C" int mvlt_window_attention(const void* qkv, void* out, int dtype, const float* relbias, int B, int H, int W,
                                     int C, int heads, int window, int shift, float scale, cudaStream_t stream) {
  if (!qkv || !out || !relbias || B <= 0 || heads <= 0 || C != heads * 32) return MVLT_ERR_INVALID;
  { const int rc0 = mvlt_attn_init(); if (rc0 != MVLT_OK) return rc0; }
  if (window != 7 || H % window || W % window || shift < 0 || shift >= window) return MVLT_ERR_UNSUPPORTED;
  AttnParams p{};
  p.qkv = qkv; p.out = out; p.ld_qkv = 3LL * C; p.ld_out = C; p.C = C; p.heads = heads; p.ntok = window * window;
  p.scale = scale; p.H = H; p.W = W; p.ws = window; p.shift = shift; p.relbias = relbias;
  dim3 grid(B * (H / window) * (W / window), heads);
  if (dtype == MVLT_BF16) {
    WinParams wp;
    wp.qkv = reinterpret_cast<const bf16*>(qkv); wp.out = reinterpret_cast<bf16*>(out);
    wp.bias_frag = reinterpret_cast<const float4*>(relbias);  // bf16 path: the fragment-layout table (see the header)
    if ((uintptr_t)relbias & 15) return MVLT_ERR_INVALID;
    wp.H = H; wp.W = W; wp.C = C; wp.heads = heads; wp.shift = shift; wp.nWh = H / window; wp.nWw = W / window;
    wp.n_items = B * wp.nWh * wp.nWw * heads; wp.scale = scale;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ctas = (wp.n_items + WA_WARPS - 1) / WA_WARPS;
    const int gridx = ctas < WA_CTAS_PER_SM * sms ? ctas : WA_CTAS_PER_SM * sms;  // 2 CTAs (98 KB, 8 warps each) per SM
    launch_k(window_attn_warp_kernel, dim3(gridx), dim3(WA_WARPS * 32), WA_WARPS * WA_WARP_BYTES, stream, wp);
  } else if (dtype == MVLT_F32) {
    const int npad = 52;
    const int bytes = (49 * 33 + 2 * npad * 33 + npad + 49 * (npad + 1)) * 4;
    launch_k(attn_f32_kernel<32, true>, dim3(grid), dim3(128), bytes, stream, p, npad, 49);
  } else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

// BERT joint attention.  qkv: [B*S, 3*768] (q|k|v), out: [B*S, 768]; kmask: [B,S] additive (ignored when seq2seq).
extern "C" int mvlt_joint_attention(const void* qkv, void* out, int dtype, const float* kmask, int B, int S, int heads,
                                    int head_dim, int seq2seq, int obj_end, float scale, cudaStream_t stream) {
  if (!qkv || !out || !kmask || B <= 0 || S <= 0 || heads <= 0) return MVLT_ERR_INVALID;
  { const int rc0 = mvlt_attn_init(); if (rc0 != MVLT_OK) return rc0; }
  if (head_dim != 64) return MVLT_ERR_UNSUPPORTED;
  const int C = heads * head_dim;
  AttnParams p{};
  p.qkv = qkv; p.out = out; p.ld_qkv = 3LL * C; p.ld_out = C; p.C = C; p.heads = heads; p.ntok = S; p.scale = scale;
  p.kmask = kmask; p.seq2seq = seq2seq; p.obj_end = obj_end;
  dim3 grid(B, heads);
  if (dtype == MVLT_BF16) {
    if (B > 65535) return MVLT_ERR_UNSUPPORTED;
    p.head_major = 1;
    grid = dim3(heads, B);
  }
  if (dtype == MVLT_BF16) {
    if (S <= 96) launch_k(joint_attn_flash_kernel<2>, dim3(grid), dim3(96), joint_flash_smem<2>(), stream, p);
    else if (S <= 144) launch_k(joint_attn_flash_kernel<3>, dim3(grid), dim3(96), joint_flash_smem<3>(), stream, p);
    else if (S <= 192) launch_k(joint_attn_flash_kernel<4>, dim3(grid), dim3(96), joint_flash_smem<4>(), stream, p);
    else if (S <= 240) launch_k(joint_attn_flash_kernel<5>, dim3(grid), dim3(96), joint_flash_smem<5>(), stream, p);   // ViT: 197 tokens
    else if (S <= 288) launch_k(joint_attn_flash_kernel<6>, dim3(grid), dim3(96), joint_flash_smem<6>(), stream, p);   // 196 image tokens + text
    else return MVLT_ERR_UNSUPPORTED;
  } else if (dtype == MVLT_F32) {
    const int npad = (S + 3) & ~3;
    auto bytes_for = [&](int rb) { return (rb * 65 + 2 * npad * 65 + npad + rb * (npad + 1)) * 4; };
    int rb = S;                                             // query rows per CTA: the whole sequence when it fits
    while (rb > 8 && bytes_for(rb) > 200 * 1024) rb = (rb + 1) / 2;
    if (bytes_for(rb) > 200 * 1024) return MVLT_ERR_UNSUPPORTED;
    launch_k(attn_f32_kernel<64, false>, dim3(B, heads, (S + rb - 1) / rb), dim3(256), bytes_for(rb), stream, p, npad, rb);
  } else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
