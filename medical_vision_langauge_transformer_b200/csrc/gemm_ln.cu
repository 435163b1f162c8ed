// Linear + bias + residual + LayerNorm as ONE tcgen05 kernel: the two post-LN sites of every BERT layer
//
//     y = LayerNorm(A[M,K] . W[768,K]^T + bias + R) * gamma + beta          (HF modeling_bert.py:295-297 BertSelfOutput,
//                                                                            :352-354 BertOutput; hidden_size = 768)
//
// replaces gemm_tc (fp32 reduce-add epilogue into the residual stream) followed by layernorm_rows: per call the fp32 [M,768]
// stream was read-modify-written by the GEMM, read again and written twice (fp32 + bf16 shadow) by the LayerNorm — 130 MB at
// M = 8384 against 77 MB here (A, R in, fp32 + bf16 out), and one launch instead of two.
//
// A LayerNorm row needs all 768 accumulator columns, a CTA's tensor memory holds 512.  So a CLUSTER OF FOUR CTAs owns 256 rows:
// CTA pair p = rank >> 1 computes the column half [384p, 384p + 384) of those rows with tcgen05.mma.cta_group::2 (M = 256, two
// N = 192 instructions per k-step; each CTA TMA-loads its own 128 rows of A and 2 x 96 rows of W per 64-wide k-block: 40 KB per
// stage, five stages), each CTA ends up with 128 rows x 384 fp32 columns in tensor memory, and the row statistics are completed
// across the two pairs through distributed shared memory (partial sums written into the partner CTA rank ^ 2, mbarrier with
// cluster-scope release/acquire).  33 such clusters are resident on a B200 (tools/micro/cluster_occ.cu) = the 33 row tiles of the
// batch-64 VQA step: one wave.
//
// Epilogue (16 warps = 4 per TMEM lane quarter, 96 columns each; one row per thread):
//   the operand ring is dead once the accumulator is complete, so the residual tile (128 x 384 fp32 = 192 KB) is TMA-loaded INTO
//   it (prefetched to L2 while the MMAs ran), then
//   pass 1: x = acc + bias + R  -> written back to tensor memory; sum of the warp's 96 columns -> its mean
//   pass 2: M2 = sum (x - that mean)^2 over tensor memory; ONE exchange of (mean, M2) per part, merged exactly (equal counts)
//   pass 3: y = (x - mean) * rstd * gamma + beta -> swizzled staging (the same ring bytes) -> TMA stores of the fp32 rows and
//           their bf16 shadow (the next GEMM's A operand).
//   Two-pass statistics (no E[x^2] - mean^2 cancellation); they differ from layernorm_rows only in summation order.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "cg2.cuh"
#include "tmap.cuh"

extern "C" int mvlt_gemm_tc_init(void);

namespace mvlt {
namespace gl {

constexpr int N = 768;                  // LayerNorm width = GEMM N
constexpr int NH = N / 2;               // columns per CTA pair (tensor-memory columns per CTA)
constexpr int SUBN = 192;               // N of one MMA instruction
constexpr int NSUB = NH / SUBN;         // 2
constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;                    // 16 KB
constexpr int B_SUB_BYTES = (SUBN / 2) * BK * 2;        // 12 KB: the W rows one CTA holds for one MMA
constexpr int STAGE_BYTES = A_BYTES + NSUB * B_SUB_BYTES;   // 40 KB
constexpr int STAGES = 5;
constexpr int RING_BYTES = STAGES * STAGE_BYTES;        // 200 KB >= the 192 KB residual tile
constexpr int EPI_WARP0 = 4, EPI_WARPS = 16, THREADS = (EPI_WARP0 + EPI_WARPS) * 32;
constexpr int PARTS = EPI_WARPS / 4;                    // warps per lane quarter
constexpr int WCOLS = NH / PARTS;                       // 96 columns per warp
constexpr int WCH = WCOLS / 32;                         // 3 chunks of 32 columns
constexpr int WARP_REGION = WCH * 4096;                 // 12 KB of the ring per epilogue warp
constexpr int PAR_BYTES = 3 * NH * 4;                   // bias, gamma, beta of this pair's columns
constexpr int STAT_SLOTS = 2 * PARTS;                   // partial sums per row: (pair, part)
constexpr int STAT_BYTES = 2 * 2 * STAT_SLOTS * BM * 4; // [tile parity][mean | M2][slot][row]
constexpr int NUM_BARS = 2 * STAGES + 4 + EPI_WARPS;
constexpr int AUX_BYTES = 512;
constexpr int SMEM_BYTES = RING_BYTES + PAR_BYTES + STAT_BYTES + AUX_BYTES + 1024;
constexpr int TMEM_COLS = 512;
static_assert(EPI_WARPS * WARP_REGION <= RING_BYTES, "the residual tile is parked in the operand ring");
static_assert(NUM_BARS * 8 + 8 <= AUX_BYTES && SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params {
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps;
  int M, K, tiles;
  int has_bf16;
  unsigned long long* trace;
};
static unsigned long long* g_trace = nullptr;
#define GL_STAMP(idx) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[(idx)] = (unsigned long long)clock64(); } while (0)

__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_o32,
               const __grid_constant__ CUtensorMap tmap_o16, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* ring = smem;                                             // stage s: A (16 KB) | W sub-tile 0 | W sub-tile 1
  float* spar = reinterpret_cast<float*>(smem + RING_BYTES);        // bias[NH] | gamma[NH] | beta[NH]
  float* sstat = reinterpret_cast<float*>(smem + RING_BYTES + PAR_BYTES);   // [2][2][STAT_SLOTS][BM]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES + PAR_BYTES + STAT_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES] TMA -> MMA (pair leader)
  uint64_t* empty_bar = bars + STAGES;            // [STAGES] MMA -> TMA (both CTAs of the pair)
  uint64_t* tmem_full = bars + 2 * STAGES;        // accumulator complete (both CTAs of the pair)
  uint64_t* tmem_empty = bars + 2 * STAGES + 1;   // epilogues of both CTAs -> MMA (pair leader)
  uint64_t* ring_free = bars + 2 * STAGES + 2;    // this CTA's epilogue warps -> its producer
  uint64_t* stat_bar = bars + 2 * STAGES + 3;     // partial row statistics of both pairs have landed
  uint64_t* res_bar = bars + 2 * STAGES + 4;      // [EPI_WARPS] residual chunks of one warp
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();        // 0..3
  const uint32_t pair = rank >> 1, r = rank & 1;  // column half; row half inside the pair (r == 0: MMA leader)
  const int cluster = blockIdx.x >> 2, num_clusters = gridDim.x >> 2;
  const int num_kb = (p.K + BK - 1) / BK;
  const int col0 = (int)pair * NH;
  if (warp == 0) GL_STAMP(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_r);
    tma_prefetch_desc(&tmap_o32); tma_prefetch_desc(&tmap_o16);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 2 * EPI_WARPS);
    mbar_init(ring_free, EPI_WARPS);
    mbar_init(stat_bar, 2 * EPI_WARPS);
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&res_bar[w], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_cg2(tmem_ptr, TMEM_COLS);
    tmem_relinquish_cg2();
  }
  if (warp >= EPI_WARP0) {     // weights: not produced by the preceding kernel, loaded ahead of the grid dependency
    const int t = threadIdx.x - EPI_WARP0 * 32;
    if (t < NH) {
      spar[t] = p.bias != nullptr ? __ldg(p.bias + col0 + t) : 0.f;
      spar[NH + t] = __ldg(p.gamma + col0 + t);
      spar[2 * NH + t] = __ldg(p.beta + col0 + t);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_grid_sync();
  if (warp == 0) GL_STAMP(1);

  if (warp == 0) {
    // ------------------------------- TMA producer (every CTA) -------------------------------
    uint32_t kc = 0, it = 0;
    for (int tile = cluster; tile < p.tiles; tile += num_clusters, ++it) {
      if (it > 0) mbar_wait(ring_free, (it - 1) & 1);     // the ring held the residual / output staging of the previous tile
      const int m0 = tile * (2 * BM) + (int)r * BM;
      for (int kb = 0; kb < num_kb; ++kb, ++kc) {
        const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          if (r == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
          uint8_t* st = ring + s * STAGE_BYTES;
          tma_load_cg2(st, &tmap_a, &full_bar[s], kb * BK, m0);
#pragma unroll
          for (int j = 0; j < NSUB; ++j)
            tma_load_cg2(st + A_BYTES + j * B_SUB_BYTES, &tmap_w, &full_bar[s], kb * BK, col0 + j * SUBN + (int)r * (SUBN / 2));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (pair leader) --------------------------------
    if (r == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM, SUBN);
      constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);   // SBO 1024 B | version 1 | SWIZZLE_128B
      const uint32_t lo0 = (base >> 4) | (1u << 16);
      const uint16_t mask = (uint16_t)(3u << (rank & 2u));
      uint32_t kc = 0, it = 0;
      for (int tile = cluster; tile < p.tiles; tile += num_clusters, ++it) {
        mbar_wait(tmem_empty, (it & 1) ^ 1);
        tc_fence_after();
        GL_STAMP(8);
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (kb == 0) GL_STAMP(9);
          const uint32_t lo_a = lo0 + s * (STAGE_BYTES >> 4);
          const uint32_t lo_b = lo_a + (A_BYTES >> 4);
          const int ksteps = min(BK, p.K - kb * BK) / 16;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (k < ksteps) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j)
                  umma_bf16_cg2(tmem_base + j * SUBN, ((uint64_t)DESC_HI << 32) | (lo_a + 2 * k),
                                ((uint64_t)DESC_HI << 32) | (lo_b + j * (B_SUB_BYTES >> 4) + 2 * k), idesc, (kb | k) != 0);
              }
            }
            umma_commit_pair(&empty_bar[s], mask);
            if (kb == num_kb - 1) umma_commit_pair(tmem_full, mask);
          }
          __syncwarp();
        }
        GL_STAMP(10);
      }
    }
  } else if (warp == 3) {
    // ------------------------------- residual tile -> L2 while the MMAs run -------------------
    uint32_t it = 0;
    for (int tile = cluster; tile < p.tiles; tile += num_clusters, ++it) {
      if (it > 0) mbar_wait(tmem_full, (it - 1) & 1);
      const int m0 = tile * (2 * BM) + (int)r * BM;
      for (int b = lane; b < 4 * (NH / 32); b += 32) {
        const int q = b / (NH / 32), c = b - q * (NH / 32);
        if (m0 + q * 32 < p.M) tma_prefetch_l2_2d(&tmap_r, col0 + c * 32, m0 + q * 32);
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------- epilogue (every CTA: 128 rows x 384 columns) -------------
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;            // TMEM lane quarter
    const int part = ew >> 2;          // columns [part * 96, +96) of this CTA's 384
    const int row = q * 32 + lane;     // row of this thread inside the CTA's 128
    uint8_t* region = ring + ew * WARP_REGION;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * WCOLS);
    const uint32_t row128 = lane * 128u, swz128 = lane & 7u;
    const uint32_t row64 = lane * 64u, swz64 = (lane >> 1) & 3u;
    const float* sb = spar + part * WCOLS;
    const float* sg = spar + NH + part * WCOLS;
    const float* sbt = spar + 2 * NH + part * WCOLS;
    const int slot = (int)pair * PARTS + part;
    // my partial (mean, M2 over my 96 columns) lands in both CTAs that hold this row: here and in the CTA of the other pair with
    // the same row half.  The buffers alternate with the tile parity: the partner can run at most one tile ahead (its exchange of
    // tile t+1 needs my arrival, which follows my reads of tile t), so buffer t & 1 is rewritten only after everyone has read it.
    float* st_local = sstat + slot * BM + row;
    const uint32_t st_remote = mapa_u32(smem_u32(st_local), rank ^ 2u);
    const uint32_t bar_remote = mapa_u32(smem_u32(stat_bar), rank ^ 2u);
    constexpr int PLANE = STAT_SLOTS * BM;     // floats per [slot][row] plane
    uint32_t it = 0;
    for (int tile = cluster; tile < p.tiles; tile += num_clusters, ++it) {
      const int m0 = tile * (2 * BM) + (int)r * BM + q * 32;
      const int n0 = col0 + part * WCOLS;
      mbar_wait(tmem_full, it & 1);
      tc_fence_after();
      if (ew == 0) GL_STAMP(16);
      if (elect_one()) {
        mbar_arrive_expect_tx(&res_bar[ew], WARP_REGION);
#pragma unroll
        for (int i = 0; i < WCH; ++i) tma_load_2d(region + i * 4096, &tmap_r, &res_bar[ew], n0 + i * 32, m0);
      }
      __syncwarp();
      mbar_wait(&res_bar[ew], it & 1);
      if (ew == 0) GL_STAMP(17);

      // ---- pass 1: x = acc + bias + residual -> tensor memory; row sum
      float sum = 0.f;
#pragma unroll 1
      for (int i = 0; i < WCH; ++i) {
        uint32_t a[32];
        tmem_ld_32x32(tmem_lane + i * 32, a);
        float4 t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = *reinterpret_cast<const float4*>(region + i * 4096 + row128 + (((uint32_t)j ^ swz128) << 4));
        tmem_ld_wait();
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(sb + i * 32 + 4 * j);
          const float x0 = __uint_as_float(a[4 * j]) + b.x + t[j].x, x1 = __uint_as_float(a[4 * j + 1]) + b.y + t[j].y;
          const float x2 = __uint_as_float(a[4 * j + 2]) + b.z + t[j].z, x3 = __uint_as_float(a[4 * j + 3]) + b.w + t[j].w;
          a[4 * j] = __float_as_uint(x0); a[4 * j + 1] = __float_as_uint(x1);
          a[4 * j + 2] = __float_as_uint(x2); a[4 * j + 3] = __float_as_uint(x3);
          s0 += x0; s1 += x1; s2 += x2; s3 += x3;
        }
        sum += (s0 + s1) + (s2 + s3);
        tmem_st_32x32(tmem_lane + i * 32, a);
      }
      tmem_st_wait();
      if (ew == 0) GL_STAMP(18);
      const float mean_w = sum * (1.0f / WCOLS);

      // ---- pass 2: squared deviations from this warp's own mean (exact two-pass statistics of its 96 columns)
      float sq = 0.f;
#pragma unroll 1
      for (int i = 0; i < WCH; ++i) {
        uint32_t a[32];
        tmem_ld_32x32(tmem_lane + i * 32, a);
        tmem_ld_wait();
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d0 = __uint_as_float(a[4 * j]) - mean_w, d1 = __uint_as_float(a[4 * j + 1]) - mean_w;
          const float d2 = __uint_as_float(a[4 * j + 2]) - mean_w, d3 = __uint_as_float(a[4 * j + 3]) - mean_w;
          s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
        }
        sq += (s0 + s1) + (s2 + s3);
      }
      if (ew == 0) GL_STAMP(19);

      // ---- ONE exchange of (mean, M2) per 96-column part; the eight parts of a row are merged with the parallel-variance
      //      formula (equal counts): mean = avg(mean_i), M2 = sum M2_i + 96 * sum (mean_i - mean)^2 — still exact two-pass
      //      arithmetic, evaluated in the same order in both pairs
      const int pb = (int)(it & 1) * 2 * PLANE;
      st_local[pb] = mean_w;
      st_local[pb + PLANE] = sq;
      st_cluster_f32(st_remote + 4u * pb, mean_w);
      st_cluster_f32(st_remote + 4u * (pb + PLANE), sq);
      fence_acq_rel_cluster();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(smem_u32(stat_bar));    // a shared::cta address is its own shared::cluster address
        mbar_arrive_cluster(bar_remote);
      }
      mbar_wait_cluster(stat_bar, it & 1);
      const float* sm = sstat + pb + row;
      float msum = 0.f, m2 = 0.f, mi[STAT_SLOTS];
#pragma unroll
      for (int i = 0; i < STAT_SLOTS; ++i) { mi[i] = sm[i * BM]; msum += mi[i]; m2 += sm[PLANE + i * BM]; }
      const float mean = msum * (1.0f / STAT_SLOTS);
      float dev = 0.f;
#pragma unroll
      for (int i = 0; i < STAT_SLOTS; ++i) dev = fmaf(mi[i] - mean, mi[i] - mean, dev);
      const float var = fmaf((float)WCOLS, dev, m2) * (1.0f / N);
      const float rstd = 1.0f / sqrtf(var + p.eps);
      if (ew == 0) GL_STAMP(21);

      // ---- pass 3: normalise, stage (fp32 SW128 box + bf16 SW64 box), TMA stores
#pragma unroll 1
      for (int i = 0; i < WCH; ++i) {
        uint32_t a[32];
        tmem_ld_32x32(tmem_lane + i * 32, a);
        uint8_t* f32buf = region + (i & 1) * 4096;
        uint8_t* b16buf = region + 8192 + (i & 1) * 2048;
        if (i >= 2) {           // buffer (i & 1) was the source of the stores of chunk i - 2
          if (elect_one()) bulk_wait_read<1>();
          __syncwarp();
        }
        tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = *reinterpret_cast<const float4*>(sg + i * 32 + 4 * j);
          const float4 b = *reinterpret_cast<const float4*>(sbt + i * 32 + 4 * j);
          y[4 * j] = fmaf((__uint_as_float(a[4 * j]) - mean) * rstd, g.x, b.x);
          y[4 * j + 1] = fmaf((__uint_as_float(a[4 * j + 1]) - mean) * rstd, g.y, b.y);
          y[4 * j + 2] = fmaf((__uint_as_float(a[4 * j + 2]) - mean) * rstd, g.z, b.z);
          y[4 * j + 3] = fmaf((__uint_as_float(a[4 * j + 3]) - mean) * rstd, g.w, b.w);
          *reinterpret_cast<float4*>(f32buf + row128 + (((uint32_t)j ^ swz128) << 4)) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        }
        if (p.has_bf16) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(b16buf + row64 + (((uint32_t)j ^ swz64) << 4)) =
                make_uint4(pack_bf16x2(y[8 * j], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                           pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (m0 < p.M) {
            tma_store_2d(&tmap_o32, f32buf, n0 + i * 32, m0);
            if (p.has_bf16) tma_store_2d(&tmap_o16, b16buf, n0 + i * 32, m0);
          }
          bulk_commit();
        }
        __syncwarp();
      }
      if (ew == 0) GL_STAMP(22);
      // accumulator columns and ring bytes of this warp are free again
      tc_fence_before();
      if (elect_one()) bulk_wait_read<0>();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_remote(tmem_empty, rank & 2u);
        mbar_arrive(ring_free);
      }
    }
    if (elect_one()) bulk_wait_all();
    __syncwarp();
    if (ew == 0) GL_STAMP(23);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, TMEM_COLS);
  }
}

static int g_max_clusters = 0;

}  // namespace gl
}  // namespace mvlt

using namespace mvlt;

// debug hook (not part of include/mvlt_b200.h): device buffer of >= 32 u64 stamped by CTA 0 of every later launch
extern "C" int mvlt_debug_gemm_ln_trace(void* dev_buf) {
  gl::g_trace = reinterpret_cast<unsigned long long*>(dev_buf);
  return MVLT_OK;
}

static int gemm_ln_configure() {
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(gl::gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gl::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
  }
  if (gl::g_max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(gl::THREADS);
    cfg.dynamicSmemBytes = gl::SMEM_BYTES;
    cfg.gridDim = dim3(4 * 32);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, gl::gemm_ln_kernel, &cfg);
    if (e != cudaSuccess) return (int)e;
    if (n <= 0) return MVLT_ERR_UNSUPPORTED;
    gl::g_max_clusters = n;
  }
  return MVLT_OK;
}

// How many 256-row tiles of mvlt_linear_residual_layernorm run at once on this device (clusters of four CTAs that can be
// resident: 33 on a 148-SM B200), or a negative / CUDA error code.  Callers use it to decide whether a given M fills a wave.
extern "C" int mvlt_linear_ln_resident_tiles(void) {
  int rc = gemm_ln_configure();
  return rc != MVLT_OK ? (rc > 0 ? -rc : rc) : gl::g_max_clusters;
}

// y = LayerNorm(A . W^T + bias + residual) (see the header of this file).  out_f32 may alias residual (every CTA reads its
// residual tile before it writes the same tile); out_bf16 may be NULL.  N must be 768, K a multiple of 16.
extern "C" int mvlt_linear_residual_layernorm(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                              const float* residual, long long ldres, const float* gamma, const float* beta,
                                              float eps, float* out_f32, long long ld32, void* out_bf16, long long ld16,
                                              int M, int N, int K, cudaStream_t stream) {
  if (!A || !W || !residual || !gamma || !beta || !out_f32 || M <= 0 || K <= 0) return MVLT_ERR_INVALID;
  if (N != gl::N) return MVLT_ERR_UNSUPPORTED;
  if (K % 16 != 0 || lda % 8 != 0 || ldw % 8 != 0 || ldres % 4 != 0 || ld32 % 4 != 0 || (out_bf16 && ld16 % 8 != 0)) return MVLT_ERR_INVALID;
  if (((uintptr_t)A & 15) || ((uintptr_t)W & 15) || ((uintptr_t)residual & 15) || ((uintptr_t)out_f32 & 15) || ((uintptr_t)out_bf16 & 15) ||
      ((uintptr_t)bias & 15) || ((uintptr_t)gamma & 15) || ((uintptr_t)beta & 15))
    return MVLT_ERR_INVALID;
  int rc = mvlt_gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  if ((rc = gemm_ln_configure()) != MVLT_OK) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(gl::THREADS);
  cfg.dynamicSmemBytes = gl::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  CUtensorMap ta, tw, tr, to32, to16;
  if ((rc = make_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, gl::BK, gl::BM, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, N, K, ldw, gl::BK, gl::SUBN / 2, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, residual, M, N, ldres, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&to32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out_f32, M, N, ld32, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE)) != MVLT_OK) return rc;
  if (out_bf16) {
    if ((rc = make_tmap(&to16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_bf16, M, N, ld16, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE)) != MVLT_OK) return rc;
  } else {
    to16 = to32;
  }
  gl::Params p;
  p.bias = bias; p.gamma = gamma; p.beta = beta; p.eps = eps; p.M = M; p.K = K;
  p.tiles = (M + 2 * gl::BM - 1) / (2 * gl::BM);
  p.has_bf16 = out_bf16 != nullptr;
  p.trace = gl::g_trace;
  const int clusters = p.tiles < gl::g_max_clusters ? p.tiles : gl::g_max_clusters;
  cfg.gridDim = dim3(4 * clusters);
  cfg.numAttrs = mvlt_pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gl::gemm_ln_kernel, ta, tw, tr, to32, to16, p);
  return e == cudaSuccess ? MVLT_OK : (int)e;
}
