// Second half of a Swin block as ONE tcgen05 kernel on CTA pairs (cta_group::2), in place on the fp32 residual stream:
//
//     x  <-  x + proj(o) + b_proj                                   vfe.py:252 (proj), :384 (shortcut + drop_path(x))
//     x  <-  x + fc2( GELU( fc1( LayerNorm2(x) ) ) )                vfe.py:385, :136-139
//
// o = the window-attention output (bf16, natural token order).  A cluster of two CTAs owns 256 token rows (128 each); one
// thread of the leader issues tcgen05.mma.cta_group::2 (M = 256), each CTA TMA-loads its own A rows and HALF of every weight
// tile, so the weight tiles cross shared memory once per pair.  The fp32 residual rows LIVE IN TENSOR MEMORY for the whole
// kernel:
//   1. MMA0: ACC2[256, C] (TMEM) = o . Wproj^T
//   2. the 16 compute warps of each CTA (one accumulator row per thread, four warps per lane quarter splitting the columns)
//      add x (TMA'd fp32 chunks) + b_proj and tcgen05.st the sum back into ACC2 — the new residual never leaves the SM —
//      and take LayerNorm statistics on the way (two passes over TMEM, partial sums exchanged through shared memory);
//      the normalised rows go to shared memory as the bf16 K-major SWIZZLE_128B A operand (A1)
//   3. hidden dimension in 128-column chunks j:  MMA1(j): ACC1 = A1 . W1[j]^T;  compute warps: + b1, erf-GELU, bf16 -> A2[j&1];
//      MMA2(j): ACC2 += A2[j&1] . W2[:, j]^T  — ACC2 already holds the residual, so the fc2 result accumulates onto it.
//      Issue order MMA1(j+1), MMA2(j): the tensor pipe works on the next chunk while the GELU of this one runs; ACC1 is
//      single-buffered (ACC2 takes C of the 512 TMEM columns) and released as soon as the GELU warps have LOADED it.
//   4. ACC2 + b2 -> shared memory -> plain TMA store to x.
// x is read once and written once (8C bytes per row), o is read once (2C); against proj GEMM (in-place) + LayerNorm + fc1 GEMM +
// fc2 GEMM (in-place) this removes two x read-modify-write round trips, the LayerNorm output, the hidden activation write +
// read and three launches per block.  with_proj = 0 drops steps 1-2's product (x <- x + MLP(LN(x)) only).
// C = 384 (stage 2: 49 row pairs of the batch-64 step) and C = 192 (stage 1).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "cg2.cuh"
#include "tmap.cuh"

extern "C" int mvlt_gemm_tc_init(void);

namespace mvlt {

constexpr int BT_THREADS = 18 * 32;        // warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-17 compute
constexpr int BT_CW0 = 2;                  // first compute warp
constexpr int BT_NCW = 16;
constexpr int BT_ACC1_COL = 384;            // ACC2 (C <= 384 columns) first, then the 128-column fc1 accumulator

// Shared-memory plan.  What the first version's clock64 trace showed (profiles/r02_tail_trace_v1.log): (1) the residual chunks
// were only requested after the proj product had retired (their buffers aliased the o tile): 8.7k cycles of exposed HBM time;
// (2) a 5-slot weight ring (60 KB) behind ~1.5k cycles of L2 latency streamed 23 B/clk where the MMAs want 32: the main loop
// ran 4.2k cycles per hidden chunk against 3.1k of tensor time; (3) LayerNorm / bias parameters fetched from global memory
// inside the passes.  Hence: x goes into TENSOR MEMORY first (the proj product then accumulates onto it) through a 3-buffer
// ring that lives where the hidden activation will (loads start at kernel entry, in parallel with the o tile); all weight
// tiles are [64 rows, 64 k] (8 KB; C-wide outputs as 128-column sub-tiles) in a deeper ring; A2 is three 16 KB half-chunk
// buffers instead of two 32 KB chunks; gamma / beta / biases are staged in shared memory.
template <int C> struct TailPlan {
  static_assert(C == 384 || C == 192 || C == 96, "Swin-S stage 0 / 1 / 2 widths");
  // k-blocks of the C-wide contractions (proj, fc1).  C = 96: the second k-block is half empty — TMA zero-fills the columns
  // past C of the o tile and of the weight tiles, the LayerNorm pass writes zeros there — and all four k-steps run
  static constexpr int KB1 = (C + 63) / 64;
  static constexpr int WN = C == 384 ? 128 : C;         // output columns per MMA of the C-wide products (proj, fc2)
  static constexpr int NSUB = C / WN;                   // such sub-tiles per k-block
  static constexpr int WROWS = WN / 2;                  // weight rows per CTA and tile
  static constexpr int SLOT = (WROWS > 64 ? WROWS : 64) * 128;   // ring slot bytes (fc1 tiles are 64 rows = 8 KB)
  static constexpr int HID = 4 * C;
  static constexpr int NCHUNK = HID / 128;
  static constexpr int XCH = C / 32;                    // 32-column fp32 chunks of a residual row
  static constexpr int A1_BYTES = KB1 * 16384;          // [128 rows][C] bf16 as KB1 k-blocks of [128][64]; first: the o tile
  static constexpr int A2_BYTES = 3 * 16384;            // three [128][64] bf16 half-chunk buffers; first: x chunk ring; last: staging
  static constexpr int NSLOT = C == 384 ? 8 : 8;
  static constexpr int PAR_BYTES = 2 * C * 4;           // gamma | beta
  static constexpr int STAT_BYTES = 2 * 128 * 4 * 4;    // per-row partial sums of the four column parts (two passes); b_proj / b2 in turn
  static constexpr int NUM_BARS = 2 * NSLOT + 16 + XCH + 3;
  static constexpr int AUX_BYTES = 768;
  static constexpr int SMEM = A1_BYTES + A2_BYTES + NSLOT * SLOT + PAR_BYTES + STAT_BYTES + AUX_BYTES + 1024;
  static_assert(NUM_BARS * 8 + 8 <= AUX_BYTES, "barrier block");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(C * 4 <= 2048, "b_proj / b2 fit one half of the statistics block");
};

struct TailParams {
  long long M;
  const float* b_proj;
  const float* gamma;
  const float* beta;
  const float* b1;
  const float* b2;
  float eps;
  int with_proj;
  unsigned long long* trace;   // debug: clock64 stamps of CTA 0 (tools/tail_trace.py); nullptr in production
};
static unsigned long long* g_tail_trace = nullptr;
#define BT_STAMP(idx) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[(idx)] = (unsigned long long)clock64(); } while (0)

template <int C>
__global__ void __launch_bounds__(BT_THREADS, 1)
swin_tail_kernel(const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_x,
                 const __grid_constant__ CUtensorMap tmap_xs, const __grid_constant__ CUtensorMap tmap_wp, const __grid_constant__ CUtensorMap tmap_w2,
                 const __grid_constant__ CUtensorMap tmap_w1, const TailParams p) {
  using P = TailPlan<C>;
  constexpr int KB1 = P::KB1, NSUB = P::NSUB, WN = P::WN, WROWS = P::WROWS, SLOT = P::SLOT, NCHUNK = P::NCHUNK, XCH = P::XCH, NSLOT = P::NSLOT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a1 = smem;
  uint8_t* a2 = a1 + P::A1_BYTES;
  uint8_t* ring = a2 + P::A2_BYTES;
  float* par = reinterpret_cast<float*>(ring + NSLOT * SLOT);         // gamma[C] | beta[C]
  float* stats = par + 2 * C;                                         // [128][4] sums | [128][4] squares (b_proj first, b2 last)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + P::STAT_BYTES);
  uint64_t* w_full = bars;                    // [NSLOT] TMA -> MMA            (leader CTA's copy counts both CTAs' bytes)
  uint64_t* w_empty = w_full + NSLOT;         // [NSLOT] MMA -> TMA            (multicast commit: one copy per CTA)
  uint64_t* a0_full = w_empty + NSLOT;        // o tile landed                  (leader)
  uint64_t* x_done = a0_full + 1;             // residual rows stored in TMEM   (leader, 2 x 16 warp arrivals)
  uint64_t* acc0_full = x_done + 1;           // MMA0 retired                   (multicast)
  uint64_t* a1_full = acc0_full + 1;          // LayerNorm rows written         (leader, 32 arrivals)
  uint64_t* acc1_full = a1_full + 1;          // MMA1(j) retired                (multicast)
  uint64_t* acc1_empty = acc1_full + 1;       // GELU warps have loaded ACC1    (leader, 32 arrivals)
  uint64_t* a2_full = acc1_empty + 1;         // [3] hidden half-chunk written  (leader, 2 x 8 warp arrivals)
  uint64_t* a2_empty = a2_full + 3;           // [3] its MMA2 k-block retired   (multicast)
  uint64_t* acc2_full = a2_empty + 3;         // last MMA2 retired              (multicast)
  uint64_t* x_full = acc2_full + 1;           // [XCH] residual chunk landed    (local)
  uint64_t* x_free = x_full + XCH;            // [3]   chunk buffer consumed    (local, 4 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + P::NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x / 2;
  const int row0 = pair * 256 + (int)rank * 128;      // this CTA's first token row
  const bool with_proj = p.with_proj != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_o); tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_xs); tma_prefetch_desc(&tmap_wp); tma_prefetch_desc(&tmap_w2);
    tma_prefetch_desc(&tmap_w1);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a0_full, 1); mbar_init(x_done, 2 * BT_NCW); mbar_init(acc0_full, 1); mbar_init(a1_full, 2 * BT_NCW);
    mbar_init(acc1_full, 1); mbar_init(acc1_empty, 2 * BT_NCW);
    for (int b = 0; b < 3; ++b) { mbar_init(&a2_full[b], BT_NCW); mbar_init(&a2_empty[b], 1); mbar_init(&x_free[b], 4); }
    mbar_init(acc2_full, 1);
    for (int c = 0; c < XCH; ++c) mbar_init(&x_full[c], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_ptr, 512);
    tmem_relinquish_cg2();
  }
  if (warp >= BT_CW0) {      // parameters (static): gamma | beta, and b_proj into the (not yet used) squares half of the statistics block
    for (int i = threadIdx.x - 64; i < C; i += BT_NCW * 32) {
      par[i] = __ldg(p.gamma + i);
      par[C + i] = __ldg(p.beta + i);
      stats[512 + i] = with_proj ? __ldg(p.b_proj + i) : 0.f;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------- TMA producer (every CTA: its rows, its half of each weight tile) -----------------
    uint32_t cnt = 0;
    auto load_w = [&](const CUtensorMap* tm, uint32_t bytes, int col, int row) {
      const uint32_t s = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
      mbar_wait(&w_empty[s], ph ^ 1);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(&w_full[s], 2 * bytes);
        tma_load_cg2(ring + s * SLOT, tm, &w_full[s], col, row);
      }
      __syncwarp();
      ++cnt;
    };
    auto load_wp = [&](int kb) { for (int h = 0; h < NSUB; ++h) load_w(&tmap_wp, WROWS * 128, kb * 64, h * WN + (int)rank * WROWS); };
    auto load_w1 = [&](int j) { for (int kb = 0; kb < KB1; ++kb) load_w(&tmap_w1, 64 * 128, kb * 64, j * 128 + (int)rank * 64); };
    auto load_w2 = [&](int j) {
      for (int kb = 0; kb < 2; ++kb)
        for (int h = 0; h < NSUB; ++h) load_w(&tmap_w2, WROWS * 128, j * 128 + kb * 64, h * WN + (int)rank * WROWS);
    };
    // weights are parameters: the ring fills while the previous kernel drains; o and x wait for it
    int kb_w = 0;
    if (with_proj)
      for (; kb_w < KB1 && (kb_w + 1) * NSUB <= NSLOT; ++kb_w) load_wp(kb_w);
    pdl_grid_sync();
    if (with_proj) {
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(a0_full, 2 * P::A1_BYTES);
        for (int kb = 0; kb < KB1; ++kb) tma_load_cg2(a1 + kb * 16384, &tmap_o, a0_full, kb * 64, row0);
      }
      __syncwarp();
    }
    // residual chunks [128 rows, 32 fp32] through the three buffers of the (still unused) A2 region
    for (int c = 0; c < XCH; ++c) {
      if (c >= 3) mbar_wait(&x_free[c % 3], ((c / 3) - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full[c], 16384);
        tma_load_2d(a2 + (c % 3) * 16384, &tmap_x, &x_full[c], c * 32, row0);
      }
      __syncwarp();
    }
    if (with_proj)
      for (; kb_w < KB1; ++kb_w) load_wp(kb_w);
    load_w1(0);
    for (int j = 0; j < NCHUNK; ++j) {
      if (j + 1 < NCHUNK) load_w1(j + 1);
      load_w2(j);
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA) ----------------------------------------------------------
    pdl_grid_sync();
    if (rank == 0) {
      const uint32_t id_w = umma_idesc_bf16(256, WN), id128 = umma_idesc_bf16(256, 128);
      constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);          // SBO 1024 B | version 1 | SWIZZLE_128B
      const uint32_t lo_a1 = (smem_u32(a1) >> 4) | (1u << 16);
      const uint32_t lo_a2 = (smem_u32(a2) >> 4) | (1u << 16);
      const uint32_t lo_w = (smem_u32(ring) >> 4) | (1u << 16);
      auto desc = [&](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
      uint32_t cnt = 0;
      // one weight tile: 4 k-steps of A (k-block at a_lo) against the tile in the ring
      auto mma_tile = [&](uint32_t d, uint32_t a_lo, uint32_t idesc, bool first) {
        const uint32_t s = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_cg2(d, desc(a_lo + 2 * k), desc(lo_w + s * (SLOT >> 4) + 2 * k), idesc, !(first && k == 0));
          umma_commit_cg2(&w_empty[s]);
        }
        __syncwarp();
        ++cnt;
      };
      BT_STAMP(0);
      if (with_proj) {
        mbar_wait(a0_full, 0);
        mbar_wait(x_done, 0);                          // ACC2 holds x + b_proj: the proj product accumulates onto it
        tc_fence_after();
        BT_STAMP(1);
        for (int kb = 0; kb < KB1; ++kb)
          for (int h = 0; h < NSUB; ++h) mma_tile(tmem_base + h * WN, lo_a1 + kb * (16384 >> 4), id_w, false);
        if (elect_one()) umma_commit_cg2(acc0_full);
        __syncwarp();
        BT_STAMP(2);
      }
      mbar_wait(a1_full, 0);
      tc_fence_after();
      BT_STAMP(3);
      auto mma1 = [&](int j) {
        if (j > 0) {                                  // the GELU warps have loaded ACC1 of chunk j - 1
          mbar_wait(acc1_empty, (j - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < KB1; ++kb) mma_tile(tmem_base + BT_ACC1_COL, lo_a1 + kb * (16384 >> 4), id128, kb == 0);
        if (elect_one()) umma_commit_cg2(acc1_full);
        __syncwarp();
      };
      auto mma2 = [&](int j) {
        for (int kb = 0; kb < 2; ++kb) {
          const int hc = 2 * j + kb, b = hc % 3;      // half-chunk index -> buffer
          mbar_wait(&a2_full[b], (hc / 3) & 1);
          tc_fence_after();
          for (int h = 0; h < NSUB; ++h) mma_tile(tmem_base + h * WN, lo_a2 + b * (16384 >> 4), id_w, false);   // ACC2 holds the residual
          if (elect_one()) {
            umma_commit_cg2(&a2_empty[b]);
            if (j == NCHUNK - 1 && kb == 1) umma_commit_cg2(acc2_full);
          }
          __syncwarp();
        }
      };
      mma1(0);
      for (int j = 0; j < NCHUNK; ++j) {
        BT_STAMP(16 + 4 * j);
        if (j + 1 < NCHUNK) mma1(j + 1);
        BT_STAMP(16 + 4 * j + 1);
        mma2(j);
        BT_STAMP(16 + 4 * j + 2);
      }
    }
  } else {
    // ------------------------------- compute warps: one accumulator row per thread ------------------------------------
    const int ew = warp - BT_CW0;
    const int quarter = warp & 3;                // TMEM lanes [32 quarter, +32)
    const int part = ew >> 2;                    // column part: chunks part, part + 4, ... of a row; 32-column slice of a hidden chunk
    const int r = quarter * 32 + lane;           // row of the CTA's tile
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t rsw = (uint32_t)r & 7u;
    float* s_sum = stats;                        // [128][4]
    float* s_sq = stats + 512;                   // [128][4]; holds b_proj until pass 2
    pdl_grid_sync();
    if (ew == 0) BT_STAMP(4);
    // x + b_proj -> ACC2 (tensor memory): the residual rows stay on chip from here to the final store
#pragma unroll 1
    for (int c = part; c < XCH; c += 4) {
      mbar_wait(&x_full[c], 0);
      const uint8_t* xb = a2 + (c % 3) * 16384 + r * 128;
      uint32_t o[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(xb + (((uint32_t)i ^ rsw) << 4));
        const float4 bp = *reinterpret_cast<const float4*>(s_sq + c * 32 + 4 * i);     // broadcast read
        o[4 * i] = __float_as_uint(v.x + bp.x); o[4 * i + 1] = __float_as_uint(v.y + bp.y);
        o[4 * i + 2] = __float_as_uint(v.z + bp.z); o[4 * i + 3] = __float_as_uint(v.w + bp.w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_free[c % 3]);
      tmem_st_32x32(tl + c * 32, o);
    }
    tmem_st_wait();
    if (with_proj) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(x_done, 0);
      mbar_wait(acc0_full, 0);
      tc_fence_after();
    }
    if (ew == 0) BT_STAMP(5);
    // pass 1: row sum of x_new = x + b_proj + o . Wproj^T (from tensor memory)
    float sum = 0.f;
#pragma unroll 1
    for (int c = part; c < XCH; c += 4) {
      uint32_t v[32];
      tmem_ld_32x32(tl + c * 32, v);
      tmem_ld_wait();
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        s0 += __uint_as_float(v[i]); s1 += __uint_as_float(v[i + 1]); s2 += __uint_as_float(v[i + 2]); s3 += __uint_as_float(v[i + 3]);
      }
      sum += (s0 + s1) + (s2 + s3);
    }
    s_sum[r * 4 + part] = sum;
    if (ew == 0) BT_STAMP(6);
    named_bar_sync(1, BT_NCW * 32);              // also: b_proj (in s_sq) is dead from here on
    const float4 ps = *reinterpret_cast<const float4*>(s_sum + r * 4);
    const float mean = ((ps.x + ps.y) + (ps.z + ps.w)) * (1.0f / (float)C);
    // pass 2: centred sum of squares
    float sq = 0.f;
#pragma unroll 1
    for (int c = part; c < XCH; c += 4) {
      uint32_t v[32];
      tmem_ld_32x32(tl + c * 32, v);
      tmem_ld_wait();
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float d0 = __uint_as_float(v[i]) - mean, d1 = __uint_as_float(v[i + 1]) - mean;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
      }
      sq += q0 + q1;
    }
    s_sq[r * 4 + part] = sq;
    if (ew == 0) BT_STAMP(7);
    named_bar_sync(1, BT_NCW * 32);
    const float4 pq = *reinterpret_cast<const float4*>(s_sq + r * 4);
    const float rstd = 1.0f / sqrtf(((pq.x + pq.y) + (pq.z + pq.w)) * (1.0f / (float)C) + p.eps);
    // pass 3: normalise -> bf16 -> A1 (K-major, 128-byte rows, 16-byte slots XOR (row & 7)); the o tile there is dead
    // (acc0_full), and so are the x chunks in the A2 region (every warp passed the barriers above)
    if (XCH % 2 != 0 && part == 3) {             // C = 96: zero the unused upper half of the last k-block (columns C .. C + 31)
      uint8_t* dst = a1 + (XCH >> 1) * 16384 + r * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dst + ((((uint32_t)(4 + i)) ^ rsw) << 4)) = make_uint4(0, 0, 0, 0);
    }
#pragma unroll 1
    for (int c = part; c < XCH; c += 4) {
      uint32_t v[32];
      tmem_ld_32x32(tl + c * 32, v);
      tmem_ld_wait();
      uint8_t* dst = a1 + (c >> 1) * 16384 + r * 128;
      const float* gm = par + c * 32;
      const float* bt = par + C + c * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 g0 = *reinterpret_cast<const float4*>(gm + 8 * i), g1 = *reinterpret_cast<const float4*>(gm + 8 * i + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bt + 8 * i), b1 = *reinterpret_cast<const float4*>(bt + 8 * i + 4);
        const float y0 = (__uint_as_float(v[8 * i]) - mean) * rstd * g0.x + b0.x, y1 = (__uint_as_float(v[8 * i + 1]) - mean) * rstd * g0.y + b0.y;
        const float y2 = (__uint_as_float(v[8 * i + 2]) - mean) * rstd * g0.z + b0.z, y3 = (__uint_as_float(v[8 * i + 3]) - mean) * rstd * g0.w + b0.w;
        const float y4 = (__uint_as_float(v[8 * i + 4]) - mean) * rstd * g1.x + b1.x, y5 = (__uint_as_float(v[8 * i + 5]) - mean) * rstd * g1.y + b1.y;
        const float y6 = (__uint_as_float(v[8 * i + 6]) - mean) * rstd * g1.z + b1.z, y7 = (__uint_as_float(v[8 * i + 7]) - mean) * rstd * g1.w + b1.w;
        *reinterpret_cast<uint4*>(dst + ((((uint32_t)((c & 1) * 4 + i)) ^ rsw) << 4)) =
            make_uint4(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3), pack_bf16x2(y4, y5), pack_bf16x2(y6, y7));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(a1_full, 0);
    if (ew == 0) BT_STAMP(8);
    // b2 for the final epilogue replaces the row sums (every warp has read them: second barrier above)
    for (int i = threadIdx.x - 64; i < C; i += BT_NCW * 32) s_sum[i] = __ldg(p.b2 + i);
    named_bar_sync(1, BT_NCW * 32);

    // main loop: GELU of hidden chunk j, this warp's 32 columns of it (half-chunk part >> 1, 16-byte slots (part & 1) * 4 ..)
    const uint32_t a2_row = (uint32_t)r * 128;
#pragma unroll 1
    for (int j = 0; j < NCHUNK; ++j) {
      const int hc = 2 * j + (part >> 1), b = hc % 3;
      mbar_wait(acc1_full, j & 1);
      tc_fence_after();
      if (ew == 0) BT_STAMP(80 + 4 * j);
      uint32_t rg[32];
      tmem_ld_32x32(tl + BT_ACC1_COL + part * 32, rg);
      const float* bias = p.b1 + j * 128 + part * 32;
      float4 bv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(bias) + i);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(acc1_empty, 0);     // ACC1 may be overwritten by MMA1(j + 1)
      float2 v[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[2 * i] = add2(make_float2(__uint_as_float(rg[4 * i]), __uint_as_float(rg[4 * i + 1])), make_float2(bv[i].x, bv[i].y));
        v[2 * i + 1] = add2(make_float2(__uint_as_float(rg[4 * i + 2]), __uint_as_float(rg[4 * i + 3])), make_float2(bv[i].z, bv[i].w));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = gelu_erf_pk2(v[i]);
      mbar_wait(&a2_empty[b], ((hc / 3) & 1) ^ 1);          // the MMA2 k-block that last read this buffer has retired
      uint8_t* dst = a2 + b * 16384 + a2_row;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(dst + ((((uint32_t)((part & 1) * 4 + i)) ^ rsw) << 4)) =
            make_uint4(pack_bf16x2(v[4 * i].x, v[4 * i].y), pack_bf16x2(v[4 * i + 1].x, v[4 * i + 1].y),
                       pack_bf16x2(v[4 * i + 2].x, v[4 * i + 2].y), pack_bf16x2(v[4 * i + 3].x, v[4 * i + 3].y));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&a2_full[b], 0);
      if (ew == 0) BT_STAMP(80 + 4 * j + 1);
    }

    // final epilogue: x = ACC2 + b2 (ACC2 = residual + fc2 product).  Each warp parks 32x32 fp32 chunks in shared memory (the A1
    // region: every MMA1 has retired) in the TMA swizzle pattern and stores them; rows past M are clipped by the TMA unit.
    mbar_wait(acc2_full, 0);
    tc_fence_after();
    if (ew == 0) BT_STAMP(9);
    uint8_t* sb = a1 + ew * 4096;
    const uint32_t srow = (uint32_t)lane * 128u, sswz = (uint32_t)lane & 7u;
#pragma unroll 1
    for (int c = part; c < XCH; c += 4) {
      uint32_t rg[32];
      tmem_ld_32x32(tl + c * 32, rg);
      if (lane == 0) bulk_wait_read<0>();      // this warp's previous store has left its staging buffer
      __syncwarp();
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bb = *reinterpret_cast<const float4*>(s_sum + c * 32 + 4 * i);
        *reinterpret_cast<float4*>(sb + srow + (((uint32_t)i ^ sswz) << 4)) =
            make_float4(__uint_as_float(rg[4 * i]) + bb.x, __uint_as_float(rg[4 * i + 1]) + bb.y,
                        __uint_as_float(rg[4 * i + 2]) + bb.z, __uint_as_float(rg[4 * i + 3]) + bb.w);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmap_xs, sb, c * 32, row0 + quarter * 32);
        bulk_commit();
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
    if (ew == 0) BT_STAMP(10);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

template <int C>
static int launch_swin_tail(const TailParams& p, const void* o, float* x, long long ldx, const void* wproj, const void* w1,
                            const void* w2, cudaStream_t stream) {
  using P = TailPlan<C>;
  static unsigned long long attr_devices = 0;
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(swin_tail_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM);
    if (e != cudaSuccess) return (int)e;
  }
  CUtensorMap to, tx, twp, tw2, tw1;
  int rc;
  if (o) {
    if ((rc = make_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, o, p.M, C, C, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  }
  if ((rc = make_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, p.M, C, ldx, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) != MVLT_OK) return rc;
  if (!o) to = tx;                               // unused when with_proj == 0 (must still be a valid encoding)
  // proj weight [C, C] and fc2 weight [C, 4C]: tiles of [WROWS rows, 64 k]; fc1 weight [4C, C]: tiles of [64 rows, 64 k]
  if ((rc = make_tmap(&twp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wproj ? wproj : w2, C, wproj ? C : P::HID, wproj ? C : P::HID, 64, P::WROWS,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tw2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w2, C, P::HID, P::HID, 64, P::WROWS, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tw1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w1, P::HID, C, C, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  // the final stores use 32-row boxes of the same x tensor
  CUtensorMap tx32;
  if ((rc = make_tmap(&tx32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, p.M, C, ldx, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE)) != MVLT_OK) return rc;
  const long long pairs = (p.M + 255) / 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(BT_THREADS);
  cfg.dynamicSmemBytes = P::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mvlt_pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, swin_tail_kernel<C>, to, tx, tx32, twp, tw2, tw1, p);
  return e == cudaSuccess ? MVLT_OK : (int)e;
}

// the persistent stage-0 variant (swin_tail96.cu)
int launch_swin_tail96(const void* o, float* x, long long ldx, const void* w_proj, const float* b_proj, const float* gamma, const float* beta,
                       float eps, const void* w1, const float* b1, const void* w2, const float* b2, long long M, cudaStream_t stream);

}  // namespace mvlt

using namespace mvlt;

// debug hook (not part of include/mvlt_b200.h): device buffer of >= 256 u64 stamped by CTA 0 of every later block-tail launch
extern "C" int mvlt_debug_tail_trace(void* dev_buf) {
  g_tail_trace = reinterpret_cast<unsigned long long*>(dev_buf);
  return MVLT_OK;
}

// x (fp32 [M, C], row stride ldx, in place) <- x + o . Wproj^T + b_proj, then x <- x + fc2(GELU(fc1(LayerNorm(x)))).
// o: bf16 [M, C] dense (the window-attention output) or NULL (then w_proj / b_proj are ignored: MLP half only).
// Weights bf16 row-major nn.Linear layout: w_proj [C, C], w1 [4C, C], w2 [C, 4C]; biases / LayerNorm parameters fp32.
// C in {192, 384}.  Replaces vfe.py:252 + :384-385 + :136-139 of one SwinTransformerBlock.
extern "C" int mvlt_swin_block_tail(const void* o, float* x, long long ldx, const void* w_proj, const float* b_proj, const float* gamma,
                                    const float* beta, float eps, const void* w1, const float* b1, const void* w2, const float* b2,
                                    long long M, int C, int hidden, cudaStream_t stream) {
  if (!x || !gamma || !beta || !w1 || !b1 || !w2 || !b2 || M <= 0) return MVLT_ERR_INVALID;
  if (o && (!w_proj || !b_proj)) return MVLT_ERR_INVALID;
  if (hidden != 4 * C || (C != 96 && C != 192 && C != 384)) return MVLT_ERR_UNSUPPORTED;
  if (ldx < C || ldx % 4 != 0 || ((uintptr_t)x & 15) || ((uintptr_t)w1 & 15) || ((uintptr_t)w2 & 15)) return MVLT_ERR_INVALID;
  if (o && (((uintptr_t)o & 15) || ((uintptr_t)w_proj & 15) || ((uintptr_t)b_proj & 15))) return MVLT_ERR_INVALID;
  if (((uintptr_t)gamma & 15) || ((uintptr_t)beta & 15) || ((uintptr_t)b1 & 15) || ((uintptr_t)b2 & 15)) return MVLT_ERR_INVALID;
  if (M > 0x7fffffffLL - 512) return MVLT_ERR_UNSUPPORTED;
  int rc = mvlt_gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  TailParams p;
  p.M = M; p.b_proj = b_proj; p.gamma = gamma; p.beta = beta; p.b1 = b1; p.b2 = b2; p.eps = eps; p.with_proj = o != nullptr;
  p.trace = g_tail_trace;
  if (C == 96) {
    // stage 0: the persistent two-tiles-in-flight kernel (swin_tail96.cu); MVLT_TAIL96=0 keeps the one-tile-per-pair kernel
    const char* e = getenv("MVLT_TAIL96");           // read per call: the tests run both kernels in one process
    const bool persistent = !(e && atoi(e) == 0);
    if (persistent) return launch_swin_tail96(o, x, ldx, w_proj, b_proj, gamma, beta, eps, w1, b1, w2, b2, M, stream);
    return launch_swin_tail<96>(p, o, x, ldx, o ? w_proj : nullptr, w1, w2, stream);
  }
  if (C == 384) return launch_swin_tail<384>(p, o, x, ldx, o ? w_proj : nullptr, w1, w2, stream);
  return launch_swin_tail<192>(p, o, x, ldx, o ? w_proj : nullptr, w1, w2, stream);
}
