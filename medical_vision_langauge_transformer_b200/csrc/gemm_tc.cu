// tcgen05 / TMEM / TMA GEMM for every nn.Linear on the MVLT hot path (bf16 operands, fp32 accumulate):
//
//     C[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual        (nn.Linear weight layout == K-major B operand)
//
// replaces: vfe.py:231 (qkv), :252 (proj), :136-139 (fc1/fc2), :443 (merge reduction);
//           HF modeling_bert.py:179-181 (Q,K,V as one [2304,768] weight), :295, :338, :352, :476, :502 (MLM decoder).
//
// One persistent, warp-specialised kernel.  A cluster of two CTAs (one SM pair) owns a 256 x BLOCK_N output tile:
// each CTA TMA-loads its own 128 rows of A and HALF of the W tile per 64-wide k-block, and one thread of the leader
// CTA issues tcgen05.mma.cta_group::2 (M=256), which reads both CTAs' smem and writes both CTAs' TMEM.
// Roles per CTA (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator,
// warps 4-11 epilogue.  Two TMEM accumulator stages (2 x 256 columns): the epilogue of tile i overlaps the MMAs of
// tile i+1.
//
// Epilogue (the part that bounds every K <= 768 call site): the kernel is compiled per
// (activation, output dtype, residual) so the inner loop carries no runtime switches.  Each epilogue warp owns
// 32 accumulator rows (its TMEM lane quarter) and walks 32-column chunks:
//   tcgen05.ld 32x32b.x32 -> + bias (packed fp32x2 adds) -> erf-GELU (packed FFMA2 polynomial + one MUFU) ->
//   [+ residual, which a TMA load has already parked in this warp's staging buffer, in place] ->
//   st.shared in the TMA swizzle pattern -> one cp.async.bulk.tensor store per chunk.
// The residual chunk for step k+1 is prefetched while step k computes (3 staging buffers per warp), M/N edges are
// clipped by the TMA unit, and no epilogue thread executes a global load/store or a bounds test per element.
// BLOCK_N is a runtime value (multiple of 32, <= 256) carried by the instruction descriptor and the TMA boxes.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "cg2.cuh"
#include "tmap.cuh"

namespace mvlt {

constexpr int CG = 2;    // CTAs per tile (cta_group::2)
constexpr int BM = 128;  // rows per CTA
constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle span
constexpr int BN_MAX = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;                // 16 KB
constexpr int B_STAGE_BYTES = (BN_MAX / CG) * BK * 2;     // 16 KB: W rows held per CTA per stage
constexpr int EPI_WARP0 = 4;
constexpr int TMEM_COLS = 512;
constexpr int AUX_BYTES = 512;
constexpr int SMEM_LIMIT = 227 * 1024;

// Shared-memory plan of one kernel variant.  The operand ring wants depth (bytes in flight per SM = L2 latency x the
// 64 B/clk the MMAs consume), the epilogue wants staging buffers; bf16 chunks are 2 KB (32 rows x 64 B), fp32 chunks
// 4 KB (32 x 128 B), and the residual variant needs a third buffer (prefetch of chunk k+1 while k-1 still drains).
// DEEP (residual GEMMs with K >= 1024: fc2 / BERT output.dense): the MMAs of a tile take >= 6000 cycles, so four
// epilogue warps with two buffers keep up and the ring gets its six stages back.
// RES: 0 no residual; 1 out-of-place residual (C = acc + R, R != C: TMA-prefetched chunks added in shared memory);
//      2 in-place residual (C += acc): the epilogue stores with cp.reduce.async.bulk.tensor .add — the fp32 add happens at
//        the L2 and the residual never enters the SM (no prefetch, no LDS, half the epilogue's shared-memory traffic, which
//        competes with the operand feed of the SS-mode MMAs for the same 128 B/clk port).
template <bool OUT_BF16, int RES, bool DEEP> struct Plan {
  // bf16 outputs (qkv, fc1/FFN-in + GELU): the per-chunk chain tcgen05.ld -> bias -> GELU -> pack -> st.shared -> fence ->
  // TMA store is latency-bound with two warps per scheduler (measured 1900 cycles per 32x32 chunk for ~300 issued
  // instructions, tools/gemm_trace.py) and those call sites are epilogue-bound: 16 warps = four per scheduler.
  // bf16 output + bf16 residual (the conv3 + identity + ReLU epilogue of a ResNet bottleneck): 16 warps, two 2 KB buffers
  static constexpr int EPI_WARPS = RES == 2 ? (DEEP ? 8 : 16) : (DEEP ? 4 : (OUT_BF16 ? 16 : 8));
  static constexpr int THREADS = (4 + EPI_WARPS) * 32;       // warps 0-3: TMA producer, MMA issuer, TMEM allocator, spare
  static constexpr int EPI_PARTS = EPI_WARPS / 4;            // warps per TMEM lane quarter
  // staging buffers per warp: the residual variant prefetches chunk k+1 while k-1 still drains (3; DEEP: 2); with four
  // warps per scheduler a single buffer is enough — its previous store drains during the next chunk's tcgen05.ld + math
  static constexpr int NBUF = RES == 1 ? ((DEEP || OUT_BF16) ? 2 : 3) : ((OUT_BF16 || RES == 2) ? 1 : 2);
  static constexpr int BUF_BYTES = OUT_BF16 ? 2048 : 4096;
  static constexpr int STAGING_BYTES = EPI_WARPS * NBUF * BUF_BYTES;
  static constexpr int EPI_BYTES = STAGING_BYTES + EPI_WARPS * 128;   // + one 32-float bias slot per warp
  static constexpr int STAGES = (SMEM_LIMIT - 1024 - AUX_BYTES - EPI_BYTES) / (A_STAGE_BYTES + B_STAGE_BYTES) > 6
                                    ? 6 : (SMEM_LIMIT - 1024 - AUX_BYTES - EPI_BYTES) / (A_STAGE_BYTES + B_STAGE_BYTES);
  static constexpr int RING_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  static constexpr int NUM_BARS = 2 * STAGES + 4 + EPI_WARPS * NBUF;
  static constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + AUX_BYTES + 1024 /*align slack*/;
  static_assert(NUM_BARS * 8 + 8 <= AUX_BYTES, "barrier block too small");
  static_assert(SMEM_BYTES <= SMEM_LIMIT && STAGES >= 4 && THREADS <= 1024, "shared memory budget");
  static_assert(!DEEP || RES != 0, "DEEP is a residual-variant plan");
};

struct GemmParams {
  const float* bias;
  int M, N, K;
  int block_n;
  int tiles_m, tiles_n;
  // implicit-GEMM convolution (conv_cpb > 0): tmap_a is an im2col-mode tensor map over the NHWC input, GEMM row m is the
  // output pixel (n, p, q) = unflatten(m; Ho*Wo, Wo) and k-block kb covers tap kb / conv_cpb = (ky, kx), channels
  // 64 * (kb % conv_cpb) .. +64 (the weight is packed tap-major: [N, R*S*C])
  int conv_cpb, conv_s, conv_wo, conv_howo, conv_stride, conv_pad;
  // ACT 5 (fused vocabulary projection + cross-entropy, model.py:396-410): instead of storing the [M, N] logits the epilogue
  // keeps a running (max, sum of exp) per row over each warp's chunks of a tile and writes ONE float2 per (row, tile, part)
  float2* lse_part;     // [M][lse_ld] (max, sum exp(x - max)) partials
  int lse_ld;           // = tiles_n * EPI_PARTS
  unsigned long long* trace;  // debug: clock64 stamps of CTA 0 (tools/gemm_trace.py); nullptr in production
};
static unsigned long long* g_gemm_trace = nullptr;
#define GEMM_STAMP(idx) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[(idx)] = (unsigned long long)clock64(); } while (0)

// ACT: 0 none, 1 erf-GELU, 2 tanh, 3 ReLU, 4 ReLU then erf-GELU.  ACT 3/4 are the ResNet epilogues: applied LAST, after the
// residual add (torchvision resnet.py Bottleneck.forward: out += identity; out = relu(out)), bf16 outputs only.
// OUT_BF16: C is bf16 (else fp32).  RES: see Plan; fp32 C takes an fp32 residual (modes 1, 2), bf16 C a bf16 residual (mode 1).
template <int ACT, bool OUT_BF16, int RES, bool DEEP>
__global__ void __launch_bounds__((Plan<OUT_BF16, RES, DEEP>::THREADS), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r, GemmParams p) {
  static_assert(!(RES == 2 && OUT_BF16), "reduce-add epilogue is fp32");
  static_assert(ACT < 3 || ACT == 5 || OUT_BF16, "ReLU epilogues are bf16-out");
  static_assert(ACT != 5 || (!OUT_BF16 && RES == 0), "the logsumexp epilogue has no output tile and no residual");
  constexpr uint32_t RES_CHUNK_BYTES = 32 * 32 * (OUT_BF16 ? 2 : 4);
  using P = Plan<OUT_BF16, RES, DEEP>;
  constexpr int STAGES = P::STAGES, EPI_NBUF = P::NBUF, EPI_BUF_BYTES = P::BUF_BYTES, RING_BYTES = P::RING_BYTES;
  constexpr int EPI_BYTES = P::EPI_BYTES, NUM_BARS = P::NUM_BARS, NUM_EPI_WARPS = P::EPI_WARPS, EPI_PARTS = P::EPI_PARTS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_epi = smem + RING_BYTES;                 // 1024 B aligned (RING_BYTES is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]  TMA -> MMA        (lives in the leader CTA)
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]  MMA -> TMA        (one copy per CTA)
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]       MMA -> epilogue   (one copy per CTA)
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]       epilogue -> MMA   (leader CTA)
  uint64_t* res_full = bars + 2 * STAGES + 4;    // [NUM_EPI_WARPS][EPI_NBUF]  residual TMA -> epilogue warp
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0) GEMM_STAMP(0);
  const uint32_t rank = cluster_ctarank();
  const int group = blockIdx.x / CG, num_groups = gridDim.x / CG;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int b_rows = p.block_n / CG;  // W rows this CTA loads per stage

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    if (RES == 1) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CG * NUM_EPI_WARPS);
    }
    for (int s = 0; s < NUM_EPI_WARPS * EPI_NBUF; ++s) mbar_init(&res_full[s], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_cg2(tmem_ptr, TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) GEMM_STAMP(1);
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // global memory is only touched below.
  pdl_grid_sync();
  if (warp == 0) GEMM_STAMP(2);

  // Producer and MMA loops run WARP-CONVERGED (all 32 lanes wait on the barriers and track the same loop state) and
  // only the tcgen05 / TMA instructions themselves are predicated on elect.sync: their operands live in uniform
  // registers, and a lane-0-only divergent region makes ptxas wrap every one of them in a vote/broadcast loop
  // (~250 issue cycles per UTCHMMA, measured) — 2x the tensor time of the instruction.
  if (warp == 0) {
    // ------------------------------- TMA producer (every CTA) -------------------------------
    const uint32_t tx_bytes = CG * (A_STAGE_BYTES + (uint32_t)b_rows * BK * 2);  // both CTAs' bytes land on the leader's barrier
    uint32_t kc = 0;
    for (int tile = group; tile < num_tiles; tile += num_groups) {
      const int m0 = (tile / p.tiles_n) * (BM * CG) + rank * BM;
      const int n0 = (tile % p.tiles_n) * p.block_n + rank * b_rows;
      // convolution: corner of the filter window of the CTA's first output pixel
      int cw = 0, ch = 0, cn = 0;
      if (p.conv_cpb > 0) {
        cn = m0 / p.conv_howo;
        const int rem = m0 - cn * p.conv_howo;
        const int op = rem / p.conv_wo;
        cw = (rem - op * p.conv_wo) * p.conv_stride - p.conv_pad;
        ch = op * p.conv_stride - p.conv_pad;
      }
      int tap = 0, cb = 0;  // filter tap and 64-channel block of k-block kb (convolution)
      for (int kb = 0; kb < num_kb; ++kb, ++kc) {
        const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          if (p.conv_cpb > 0) {
            const int ky = tap / p.conv_s;
            tma_load_im2col_cg2(smem_a + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], cb * BK, cw, ch, cn,
                                (uint16_t)(tap - ky * p.conv_s), (uint16_t)ky);
          } else {
            tma_load_cg2(smem_a + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BK, m0);
          }
          tma_load_cg2(smem_b + s * B_STAGE_BYTES, &tmap_b, &full_bar[s], kb * BK, n0);
        }
        __syncwarp();
        if (++cb == p.conv_cpb) { cb = 0; ++tap; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA) --------------------------------
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM * CG, (uint32_t)p.block_n);
      // smem matrix descriptor, K-major SW128: hi word = SBO 1024 B | version 1 | SWIZZLE_128B; lo word = addr>>4 | LBO 1
      constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);
      const uint32_t lo_a0 = (base >> 4) | (1u << 16);
      const uint32_t lo_b0 = ((base + STAGES * A_STAGE_BYTES) >> 4) | (1u << 16);
      uint32_t kc = 0, it = 0;
      for (int tile = group; tile < num_tiles; tile += num_groups, ++it) {
        const uint32_t acc = it & 1, acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        if (it < 16) GEMM_STAMP(16 + 4 * it);
        const uint32_t tmem_d = tmem_base + acc * BN_MAX;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (kb == 0 && it < 16) GEMM_STAMP(16 + 4 * it + 1);
          const uint32_t lo_a = lo_a0 + s * (A_STAGE_BYTES >> 4);
          const uint32_t lo_b = lo_b0 + s * (B_STAGE_BYTES >> 4);
          const int ksteps = min(BK, p.K - kb * BK) / 16;  // K % 16 == 0 is checked on the host
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements = 32 B inside the 128 B swizzle span: +2 in the (addr >> 4) field
              if (k < ksteps)
                umma_bf16_cg2(tmem_d, ((uint64_t)DESC_HI << 32) | (lo_a + 2 * k), ((uint64_t)DESC_HI << 32) | (lo_b + 2 * k),
                              idesc, (kb | k) != 0);
            }
            umma_commit_cg2(&empty_bar[s]);  // smem slot reusable (in both CTAs) once these MMAs retire
            if (kb == num_kb - 1) umma_commit_cg2(&tmem_full[acc]);  // accumulator complete (both CTAs' epilogues)
          }
          __syncwarp();
          if (kb == num_kb - 1 && it < 16) GEMM_STAMP(16 + 4 * it + 2);
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + NUM_EPI_WARPS) {
    // ------------------------------- epilogue (every CTA, its own 128 rows) -----------------
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;   // TMEM lanes [32*quarter, +32) are the only ones this warp may touch
    const int part = ew >> 2;       // this warp takes the chunks g == part (mod EPI_PARTS) of the CTA's chunk sequence
    const int chunks = p.block_n >> 5;
    const int my_tiles = (num_tiles - group + num_groups - 1) / num_groups;
    uint8_t* bufs = smem_epi + ew * (EPI_NBUF * EPI_BUF_BYTES);
    float* sbias = reinterpret_cast<float*>(smem_epi + P::STAGING_BYTES) + ew * 32;  // this warp's bias slot (32 columns)
    uint64_t* rbar = res_full + ew * EPI_NBUF;
    // byte offset of 16 B slot j of this thread's row inside a staging buffer, TMA swizzle applied:
    //   fp32: 128 B rows, SWIZZLE_128B: slot j -> j ^ (row & 7);   bf16: 64 B rows, SWIZZLE_64B: slot j -> j ^ ((row >> 1) & 3)
    const uint32_t row_base = OUT_BF16 ? lane * 64u : lane * 128u;
    const uint32_t swz = OUT_BF16 ? ((lane >> 1) & 3u) : (lane & 7u);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);

    // Position of this warp in its chunk sequence g = part, part + EPI_PARTS, ... (g = ti * chunks + c), advanced
    // incrementally: the divisions by runtime tile counts happen once per tile, not three times per chunk (they were a
    // third of the instructions the issue-bound epilogue executed, ncu source page of v8).
    struct Pos { int ti, c, m0, n0t; };  // tile iteration, chunk in tile, row0 of this warp's 32 rows, column0 of the tile
    auto tile_origin = [&](Pos& s) {
      const int tile = group + s.ti * num_groups;
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      s.m0 = tm * (BM * CG) + rank * BM + quarter * 32;
      s.n0t = tn * p.block_n;
    };
    auto normalize = [&](Pos& s) {
      if (s.c >= chunks) {
        do { s.c -= chunks; ++s.ti; } while (s.c >= chunks);
        if (s.ti < my_tiles) tile_origin(s);
      }
    };
    // bias of a chunk, one column per lane (0 beyond N: no ragged-edge path in the chunk loop).  Fetched one step ahead
    // and parked in shared memory: the chunk loop never waits on an L2 round trip for it.
    auto bias_fetch = [&](int n0) -> float {
      const int n = n0 + lane;
      return (p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f;
    };
    Pos cur;
    cur.ti = 0; cur.c = part; cur.m0 = 0; cur.n0t = 0;
    tile_origin(cur);
    normalize(cur);
    float bias_next = cur.ti < my_tiles ? bias_fetch(cur.n0t + cur.c * 32) : 0.f;
    if (RES == 1) {
      if (cur.ti < my_tiles) {
        const int m0 = cur.m0, n0 = cur.n0t + cur.c * 32;
        if (n0 < p.N && m0 < p.M && elect_one()) {
          mbar_arrive_expect_tx(&rbar[0], RES_CHUNK_BYTES);
          tma_load_2d(bufs, &tmap_r, &rbar[0], n0, m0);
        }
        __syncwarp();
      }
    }
    int cur_ti = -1;  // tile iteration whose accumulator this warp currently holds
    auto release_tile = [&](int ti) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[ti & 1], 0);
      if (ew == 0 && ti < 16) GEMM_STAMP(128 + 4 * ti + 1);
    };
    auto acquire_tile = [&](int ti) {
      mbar_wait(&tmem_full[ti & 1], (ti >> 1) & 1);
      tc_fence_after();
      if (ew == 0 && ti < 16) GEMM_STAMP(128 + 4 * ti);
    };
    int k = 0;
    uint32_t res_parity = 0;  // bit b: parity of the next residual load to land in staging buffer b
    float lse_m = -INFINITY, lse_s = 0.f;   // ACT 5: running row maximum / sum of exp over this warp's chunks of the tile
    for (; cur.ti < my_tiles; ++k) {
      const int ti = cur.ti, c = cur.c;
      const int buf = k % EPI_NBUF;
      uint8_t* sb = bufs + buf * EPI_BUF_BYTES;
      const int m0 = cur.m0, n0 = cur.n0t + c * 32;
      const bool live = n0 < p.N && m0 < p.M;  // warp-uniform: the chunk holds at least one real element
      Pos nxt = cur;
      nxt.c += EPI_PARTS;
      normalize(nxt);
      const bool has_next = nxt.ti < my_tiles;
      const int m1 = nxt.m0, n1 = nxt.n0t + nxt.c * 32;
      const bool tr_on = ew == 0 && k < 24;
      if (tr_on) GEMM_STAMP(256 + 8 * k);
      sbias[lane] = bias_next;  // visible to the whole warp after the __syncwarp below
      if (has_next) bias_next = bias_fetch(n1);
      // at most the store of step k-1 is still reading shared memory.  3 buffers (RES): buffer (k+1) % 3, last used by
      // step k-2, is free for the next residual chunk; 2 buffers: buffer k % 2 (step k-2) is free for this step's writes
      if (RES == 1 && elect_one()) {
        // at most the store of step k-1 is still reading shared memory: buffer (k+1) % 3, last used by step k-2, is free
        // for the next residual chunk.  DEEP (2 buffers): the prefetch target was step k-1's buffer -> drain everything
        if (EPI_NBUF == 2) bulk_wait_read<0>(); else bulk_wait_read<1>();
        if (has_next) {
          if (n1 < p.N && m1 < p.M) {
            const int nb = (k + 1) % EPI_NBUF;
            mbar_arrive_expect_tx(&rbar[nb], RES_CHUNK_BYTES);
            tma_load_2d(bufs + nb * EPI_BUF_BYTES, &tmap_r, &rbar[nb], n1, m1);
          }
        }
      }
      __syncwarp();
      if (tr_on) GEMM_STAMP(256 + 8 * k + 1);
      while (cur_ti < ti) {  // also passes through tiles in which this warp owns no chunk (block_n = 32)
        if (cur_ti >= 0) release_tile(cur_ti);
        ++cur_ti;
        acquire_tile(cur_ti);
      }
      if (tr_on) GEMM_STAMP(256 + 8 * k + 2);
      if (live) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_lane + (uint32_t)((ti & 1) * BN_MAX + c * 32), r);
        if (RES == 1) {
          mbar_wait(&rbar[buf], (res_parity >> buf) & 1);
          res_parity ^= 1u << buf;
        }
        tmem_ld_wait();
        if (tr_on) GEMM_STAMP(256 + 8 * k + 3);
        float2 v[16];  // this thread's row: 32 consecutive columns as 16 packed pairs
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = *(reinterpret_cast<const float4*>(sbias) + j);  // broadcast read
            v[2 * j] = add2(v[2 * j], make_float2(b.x, b.y));
            v[2 * j + 1] = add2(v[2 * j + 1], make_float2(b.z, b.w));
          }
        }
        if (ACT == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = gelu_erf_pk2(v[i]);
        } else if (ACT == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = make_float2(tanhf(v[i].x), tanhf(v[i].y));
        } else if ((ACT == 3 || ACT == 4) && RES == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = make_float2(fmaxf(v[i].x, 0.f), fmaxf(v[i].y, 0.f));
            if (ACT == 4) v[i] = gelu_erf_pk2(v[i]);
          }
        }
        if (tr_on) GEMM_STAMP(256 + 8 * k + 4);
        if (ACT == 5) {
          // online logsumexp over the chunk's (valid) columns; columns past N carry zero-filled operands: excluded
          const int ncol = p.N - n0;
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (2 * i < ncol) cm = fmaxf(cm, v[i].x);
            if (2 * i + 1 < ncol) cm = fmaxf(cm, v[i].y);
          }
          const float L2E = 1.4426950408889634f, nm = -cm * L2E;
          float cs = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float e0, e1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(v[i].x, L2E, nm)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(v[i].y, L2E, nm)));
            if (2 * i < ncol) cs += e0;
            if (2 * i + 1 < ncol) cs += e1;
          }
          const float nmx = fmaxf(lse_m, cm);
          float a0, a1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a0) : "f"((lse_m - nmx) * L2E));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a1) : "f"((cm - nmx) * L2E));
          lse_s = lse_s * a0 + cs * a1;
          lse_m = nmx;
        }
        if (ACT != 5) {
        if (RES != 1) {  // buffer k % NBUF was last read by the store of step k - NBUF
          if (elect_one()) bulk_wait_read<EPI_NBUF - 1>();
          __syncwarp();
        }
        if (OUT_BF16) {
          if (RES == 1) {  // bf16 residual chunk parked in this buffer by TMA: add in fp32, then the ReLU epilogues
            uint4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = *reinterpret_cast<const uint4*>(sb + row_base + (((uint32_t)j ^ swz) << 4));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[4 * j] = add2(v[4 * j], unpack_bf16x2(t[j].x));
              v[4 * j + 1] = add2(v[4 * j + 1], unpack_bf16x2(t[j].y));
              v[4 * j + 2] = add2(v[4 * j + 2], unpack_bf16x2(t[j].z));
              v[4 * j + 3] = add2(v[4 * j + 3], unpack_bf16x2(t[j].w));
            }
            if (ACT >= 3) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] = make_float2(fmaxf(v[i].x, 0.f), fmaxf(v[i].y, 0.f));
                if (ACT == 4) v[i] = gelu_erf_pk2(v[i]);
              }
            }
          }
          // 8 bf16 (4 pairs) per 16 B slot
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(sb + row_base + (((uint32_t)j ^ swz) << 4)) =
                make_uint4(pack_bf16x2(v[4 * j].x, v[4 * j].y), pack_bf16x2(v[4 * j + 1].x, v[4 * j + 1].y),
                           pack_bf16x2(v[4 * j + 2].x, v[4 * j + 2].y), pack_bf16x2(v[4 * j + 3].x, v[4 * j + 3].y));
        } else {
          float4 t[8];
          if (RES == 1) {  // all eight loads before the first store: the in-place slots alias as far as the compiler can tell
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = *reinterpret_cast<const float4*>(sb + row_base + (((uint32_t)j ^ swz) << 4));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float2 a = v[2 * j], b = v[2 * j + 1];
            if (RES == 1) {
              a = add2(a, make_float2(t[j].x, t[j].y));
              b = add2(b, make_float2(t[j].z, t[j].w));
            }
            *reinterpret_cast<float4*>(sb + row_base + (((uint32_t)j ^ swz) << 4)) = make_float4(a.x, a.y, b.x, b.y);
          }
        }
        if (tr_on) GEMM_STAMP(256 + 8 * k + 5);
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA unit
        }  // ACT != 5
      }
      __syncwarp();
      if (tr_on) GEMM_STAMP(256 + 8 * k + 6);
      if (ACT == 5) {
        // last chunk of this warp in the tile: one (max, sum) pair per row and (tile column, part)
        if (!has_next || nxt.ti != ti) {
          const int tile = group + ti * num_groups;
          const int tn = tile % p.tiles_n;
          const int row = m0 + lane;
          if (row < p.M) p.lse_part[(long long)row * p.lse_ld + tn * EPI_PARTS + part] = make_float2(lse_m, lse_s);
          lse_m = -INFINITY; lse_s = 0.f;
        }
      } else if (elect_one()) {
        if (live) {
          if (RES == 2) tma_reduce_add_2d(&tmap_c, sb, n0, m0); else tma_store_2d(&tmap_c, sb, n0, m0);
        }
        bulk_commit();  // one group per step (empty when the chunk lies beyond N) keeps the wait_group arithmetic uniform
      }
      __syncwarp();
      if (tr_on) GEMM_STAMP(256 + 8 * k + 7);
      cur = nxt;
    }
    while (cur_ti < my_tiles - 1) {
      if (cur_ti >= 0) release_tile(cur_ti);
      ++cur_ti;
      acquire_tile(cur_ti);
    }
    if (cur_ti >= 0) release_tile(cur_ti);
    if (elect_one()) bulk_wait_all();  // all of this warp's stores have left shared memory and are globally performed
    __syncwarp();
    if (ew == 0) GEMM_STAMP(3);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) GEMM_STAMP(4);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int g_num_sms = 0;

typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, GemmParams);

struct Variant {
  GemmKernel kernel;
  int smem, threads;
};
template <int ACT, bool OUT_BF16, int RES, bool DEEP> static Variant variant() {
  using P = Plan<OUT_BF16, RES, DEEP>;
  return Variant{gemm_tc_kernel<ACT, OUT_BF16, RES, DEEP>, P::SMEM_BYTES, P::THREADS};
}
// act 0..2: res 0..2 with fp32 C, res 0 with bf16 C; deep only with res != 0 and act == 0.  act 3/4 (ReLU epilogues): bf16 C
// with res 0 or 1 (bf16 residual).  Returns a null kernel for combinations that are not compiled.
static Variant pick_variant(int act, bool out_bf16, int res, bool deep) {
  if (act == 5) return (!out_bf16 && res == 0 && !deep) ? variant<5, false, 0, false>() : Variant{nullptr, 0, 0};
#define MVLT_ACT(O, R, D) (act == 0 ? variant<0, O, R, D>() : act == 1 ? variant<1, O, R, D>() : variant<2, O, R, D>())
  if (act >= 3) {
    if (!out_bf16 || res == 2) return Variant{nullptr, 0, 0};
    if (res == 1) return act == 3 ? variant<3, true, 1, false>() : variant<4, true, 1, false>();
    return act == 3 ? variant<3, true, 0, false>() : variant<4, true, 0, false>();
  }
  if (out_bf16) return res == 0 ? MVLT_ACT(true, 0, false) : (res == 1 && act == 0 ? variant<0, true, 1, false>() : Variant{nullptr, 0, 0});
  if (res == 1) return deep ? variant<0, false, 1, true>() : MVLT_ACT(false, 1, false);
  if (res == 2) return deep ? variant<0, false, 2, true>() : MVLT_ACT(false, 2, false);
  return MVLT_ACT(false, 0, false);
#undef MVLT_ACT
}

static PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

static int gemm_tc_init() {
  static unsigned long long devices = 0;
  if (!first_use_on_device(devices)) return MVLT_OK;
  if (!g_encode) {
    void* fn = nullptr;
    void* fn2 = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return MVLT_ERR_DRIVER;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn2, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn2) return MVLT_ERR_DRIVER;
    g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn2);
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  for (int act = 0; act < 6; ++act)
    for (int o = 0; o < 2; ++o)
      for (int r = 0; r < 3; ++r)
        for (int d = 0; d < 2; ++d) {
          if (d && (!r || act != 0 || o)) continue;
          const Variant v = pick_variant(act, o != 0, r, d != 0);
          if (!v.kernel) continue;
          cudaError_t e = cudaFuncSetAttribute(v.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem);
          if (e != cudaSuccess) return (int)e;
        }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  return MVLT_OK;
}

// 2-D row-major tensor map: dims {cols, rows}, row stride ld_elems, box {box_cols, box_rows}  (declared in tmap.cuh)
int make_tmap(CUtensorMap* map, CUtensorMapDataType dt, int elt_bytes, const void* ptr, long long rows, long long cols,
                     long long ld_elems, int box_cols, int box_rows, CUtensorMapSwizzle swz, CUtensorMapL2promotion promo) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * elt_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVLT_OK : MVLT_ERR_DRIVER;
}

// Tile-width heuristic.  Per 64-wide k-block a CTA moves 16 KB of A and 64*BN bytes of W through shared memory twice
// (TMA fill + tensor-core read) for 2*BN tensor cycles, so the operand feed costs (16384/BN + 64) B per tensor cycle:
// wide tiles win even when they quantise worse over the 74 CTA pairs (measured: profiles/r01_gemm_sweep_v3.log).
static int pick_block_n(int M, int N, int K, int groups) {
  const int cand[] = {256, 192, 128, 96, 64, 32};
  const int tiles_m = (M + BM * CG - 1) / (BM * CG);
  double best = -1;
  int best_bn = 128;
  for (int bn : cand) {
    const int tn = (N + bn - 1) / bn;
    const double useful = (double)N / ((double)tn * bn);
    const long long tiles = (long long)tiles_m * tn;
    const long long waves = (tiles + groups - 1) / groups;
    const double wave_eff = (double)tiles / ((double)waves * groups);
    // short-K call sites are bound by their epilogue / HBM traffic, not by the operand feed: discount it
    const double feed = pow(128.0 / (16384.0 / bn + 64.0), K >= 768 ? 1.0 : K / 768.0);
    const double score = useful * wave_eff * feed;
    if (score > best + 1e-9) {
      best = score;
      best_bn = bn;
    }
  }
  return best_bn;
}

}  // namespace mvlt

using namespace mvlt;

extern "C" int mvlt_gemm_tc_init(void) { return gemm_tc_init(); }

// debug hook (not part of include/mvlt_b200.h): device buffer of >= 256 u64 stamped by CTA 0 of every later launch
extern "C" int mvlt_debug_gemm_trace(void* dev_buf) {
  g_gemm_trace = reinterpret_cast<unsigned long long*>(dev_buf);
  return MVLT_OK;
}

// Geometry of an implicit-GEMM convolution over an NHWC bf16 activation (nullptr = plain GEMM)
struct ConvGeom {
  int B, H, W, C, R, S, stride, pad, Ho, Wo;
};

// im2col-mode map over x[B, H, W, C] (bf16): 64 channels x 128 output pixels per load (CUTLASS fprop conventions:
// lower corner = -pad, upper corner = pad - (filter - 1), traversal stride = convolution stride)
static int make_tmap_im2col(CUtensorMap* map, const void* x, const ConvGeom& g) {
  cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.W * g.C * 2, (cuuint64_t)g.H * g.W * g.C * 2};
  int lower[2] = {-g.pad, -g.pad};
  int upper[2] = {g.pad - (g.S - 1), g.pad - (g.R - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, lower, upper,
                               BK, BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVLT_OK : MVLT_ERR_DRIVER;
}

static int gemm_launch(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc,
                       const float* bias, const void* residual, long long ldres, int res_dtype, int M, int N, int K,
                       int act, int out_dtype, int block_n, const ConvGeom* conv, cudaStream_t stream, float2* lse_part = nullptr) {
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0) return MVLT_ERR_INVALID;
  if (K % 16 != 0 || ldw % 8 != 0 || (!conv && lda % 8 != 0)) return MVLT_ERR_INVALID;  // TMA: 16 B aligned rows
  if (((uintptr_t)A & 15) || ((uintptr_t)W & 15) || ((uintptr_t)C & 15)) return MVLT_ERR_INVALID;
  if (act < 0 || act > 5 || (act == 5) != (lse_part != nullptr) || (out_dtype != MVLT_F32 && out_dtype != MVLT_BF16)) return MVLT_ERR_INVALID;
  const bool out_bf16 = out_dtype == MVLT_BF16;
  if (ldc % (out_bf16 ? 8 : 4) != 0) return MVLT_ERR_INVALID;                // TMA store: 16 B aligned rows
  if (bias && ((uintptr_t)bias & 15)) return MVLT_ERR_INVALID;
  const bool res = residual != nullptr;
  if (res) {
    // fp32 C: the fp32 residual stream of the model (Swin proj/fc2 in place, BERT attention/FFN outputs);
    // bf16 C: the bf16 identity branch of a ResNet bottleneck
    if (res_dtype != (out_bf16 ? MVLT_BF16 : MVLT_F32)) return MVLT_ERR_UNSUPPORTED;
    if (ldres % (out_bf16 ? 8 : 4) != 0 || ((uintptr_t)residual & 15)) return MVLT_ERR_INVALID;
  }
  int rc = gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  const int groups = g_num_sms / CG;
  if (block_n <= 0) block_n = pick_block_n(M, N, K, groups);
  if (block_n % 32 != 0 || block_n > BN_MAX) return MVLT_ERR_INVALID;

  const bool deep = res && act == 0 && K >= 1024 && !out_bf16;
  // in place (C is the fp32 residual): reduce-add stores; anything else: TMA-prefetched residual chunks
  const int res_mode = !res ? 0 : ((!out_bf16 && residual == C && ldres == ldc) ? 2 : 1);
  const Variant var = pick_variant(act, out_bf16, res_mode, deep);
  if (!var.kernel) return MVLT_ERR_UNSUPPORTED;

  CUtensorMap ta, tb, tc, tr;
  const CUtensorMapDataType cdt = out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle cswz = out_bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (conv) {
    if ((rc = make_tmap_im2col(&ta, A, *conv)) != MVLT_OK) return rc;
  } else if ((rc = make_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, N, K, ldw, BK, block_n / CG, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tc, cdt, out_bf16 ? 2 : 4, C, M, N, ldc, 32, 32, cswz, CU_TENSOR_MAP_L2_PROMOTION_NONE)) != MVLT_OK) return rc;
  if (res) {
    if ((rc = make_tmap(&tr, cdt, out_bf16 ? 2 : 4, residual, M, N, ldres, 32, 32, cswz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) != MVLT_OK) return rc;
  } else {
    tr = tc;
  }

  GemmParams p;
  p.bias = bias; p.M = M; p.N = N; p.K = K; p.block_n = block_n;
  p.tiles_m = (M + BM * CG - 1) / (BM * CG);
  p.tiles_n = (N + block_n - 1) / block_n;
  p.conv_cpb = 0; p.conv_s = 1; p.conv_wo = 1; p.conv_howo = 1; p.conv_stride = 1; p.conv_pad = 0;
  if (conv) {
    p.conv_cpb = conv->C / BK; p.conv_s = conv->S; p.conv_wo = conv->Wo; p.conv_howo = conv->Ho * conv->Wo;
    p.conv_stride = conv->stride; p.conv_pad = conv->pad;
  }
  p.trace = g_gemm_trace;
  p.lse_part = lse_part;
  p.lse_ld = p.tiles_n * 2;   // Plan<false, 0, false>::EPI_PARTS
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = CG * (tiles < groups ? tiles : groups);

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(var.threads);
  cfg.dynamicSmemBytes = var.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mvlt_pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, var.kernel, ta, tb, tc, tr, p);
  return e == cudaSuccess ? MVLT_OK : (int)e;
}

extern "C" int mvlt_gemm_bf16_tc(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc,
                                 const float* bias, const void* residual, long long ldres, int res_dtype, int M, int N,
                                 int K, int act, int out_dtype, int block_n, cudaStream_t stream) {
  return gemm_launch(A, lda, W, ldw, C, ldc, bias, residual, ldres, res_dtype, M, N, K, act, out_dtype, block_n, nullptr,
                     stream);
}

extern "C" int mvlt_conv2d_nhwc_bf16_tc(const void* x, int B, int H, int W, int C, const void* w, long long ldw, void* out,
                                        long long ldc, const float* bias, const void* residual, long long ldres, int N,
                                        int R, int S, int stride, int pad, int act, int block_n, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || R <= 0 || S <= 0 || stride <= 0 || stride > 8 || pad < 0) return MVLT_ERR_INVALID;
  if (C % BK != 0) return MVLT_ERR_UNSUPPORTED;                    // one k-block = 64 channels of one filter tap
  if (pad > 127 || R - 1 - pad > 128 || S - 1 - pad > 128) return MVLT_ERR_UNSUPPORTED;  // im2col corner range (rank 4)
  ConvGeom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.R = R; g.S = S; g.stride = stride; g.pad = pad;
  g.Ho = (H + 2 * pad - R) / stride + 1;
  g.Wo = (W + 2 * pad - S) / stride + 1;
  if (g.Ho <= 0 || g.Wo <= 0) return MVLT_ERR_INVALID;
  const long long M = (long long)B * g.Ho * g.Wo;
  if (M > 0x7fffffffLL) return MVLT_ERR_UNSUPPORTED;
  return gemm_launch(x, 0, w, ldw, out, ldc, bias, residual, ldres, MVLT_BF16, (int)M, N, R * S * C, act, MVLT_BF16, block_n, &g,
                     stream);
}

// Fused vocabulary projection + logsumexp (the MLM decoder of HF modeling_bert.py:502-512 feeding the cross-entropy of model.py:404-410):
// part[m][t * 2 + e] = (max, sum exp(x - max)) of the logits A[m,:] . W[n,:]^T + bias[n] over the columns of tile t handled by epilogue
// part e — the [M, N] logits (312 MB fp32 at batch 32) are never written.  ld_part >= 2 * ceil(N / 256) float2 per row.
extern "C" int mvlt_gemm_bf16_lse_partials(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* part,
                                           long long ld_part, int M, int N, int K, cudaStream_t stream) {
  if (!part || ((uintptr_t)part & 15)) return MVLT_ERR_INVALID;
  const int tiles_n = (N + 255) / 256;
  if (ld_part != 2LL * tiles_n) return MVLT_ERR_INVALID;
  // the (unused) C tensor map is encoded over the partials buffer viewed as fp32 [M, 2 * ld_part]
  return gemm_launch(A, lda, W, ldw, part, 2 * ld_part, bias, nullptr, 0, -1, M, N, K, 5, MVLT_F32, 256, nullptr, stream,
                     reinterpret_cast<float2*>(part));
}
