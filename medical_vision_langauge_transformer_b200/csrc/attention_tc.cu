// Attention cores of the MVLT forward on the 5th-generation tensor cores (tcgen05 / TMEM, operands by TMA):
//
//   mvlt_window_attention_tc   vfe.py:234-251 inside SwinTransformerBlock.forward :360-381 — per (image, 7x7 window, head):
//                              softmax(scale q.k^T + rel_pos_bias[h] (+ -100 shift mask)) . v
//   mvlt_joint_attention_tc    HF modeling_bert.py:115-140 (eager attention) with the masks of model.py:118-128, :162-183 —
//                              per (sample, head): softmax(q.k^T / 8 + key mask | seq2seq mask) . v
//
// Both kernels share one skeleton (one persistent CTA per SM):
//   warp 0          TMA producer: Q / K / V boxes of a tile -> the tile's shared-memory slot (SWIZZLE_64B rows of 32 bf16 for
//                   head_dim 32, SWIZZLE_128B rows of 64 for head_dim 64)
//   warp 1          MMA issuer (one elected lane): S = Q.K^T into the slot's TMEM columns, later O = P.V with P read FROM
//                   TENSOR MEMORY (A operand in TMEM) and V as an MN-major B operand straight from the TMA'd rows; it polls the
//                   barriers of all slots and serves whichever is ready
//   warps 2..       one warpgroup (4 warps = 128 accumulator lanes) per slot: ONE SCORE ROW PER THREAD — tcgen05.ld the row,
//                   bias / mask add, max, exp2, sum (no shuffles, no shared memory), P as bf16 pairs back into TMEM
//                   (tcgen05.st, aliasing the dead score columns), then after the second MMA tcgen05.ld the output row,
//                   scale by 1/sum and store 64 / 128 contiguous bytes.  S and P never leave the SM.
//
// Window kernel: the token rows arrive WINDOW-MAJOR (row (b*nW + w)*49 + i, written in that order by the LayerNorm in front
// of the qkv GEMM, mvlt_layernorm_rows_winmajor), so torch.roll + window_partition (vfe.py:144-156, :361) cost nothing and a
// window is one 49-row TMA box per matrix.  A 128-lane tile holds TWO windows (lanes 0..48 and 64..112); S is the 128x128
// product whose diagonal 64x64 blocks are used; O = P[128 x 64 local keys] . [V_a | V_b] (N = 64, two MN-major atoms).  The
// output is scattered to natural token order (= window_reverse + reverse roll, vfe.py:159-173, :378).  A CTA serves ONE
// head, so that head's (rel-pos bias + shift mask) * log2e table [class][49][52] stays in shared memory.
//
// Joint kernel: a tile is 128 CONSECUTIVE rows of the [B*S, 3*768] qkv matrix for one head; it touches up to NS samples, whose
// K / V are loaded into separate buffers and multiplied under the instruction's DISABLE-OUTPUT-LANE mask so that every lane
// gets the scores of its own sample in the same TMEM columns — every lane is a real row, no padding of S = 131 to 256.
//
// SASS: UTCHMMA (SS and TS forms), LDTM / STTM, UTMALDG.  Hardware checks of every descriptor form: tools/micro/umma_probe.cu.
#include <cuda.h>

#include "common.cuh"
#include "tmap.cuh"

extern "C" int mvlt_gemm_tc_init(void);

namespace mvlt {

constexpr float AT_LOG2E = 1.4426950408889634f;
constexpr float AT_NEG_BIG = -1.0e30f;
constexpr long long AT_WATCHDOG = 4000000000LL;   // cycles without progress before the MMA scheduler traps

__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// =====================================================================================================================
// Swin window attention
// =====================================================================================================================
constexpr int WT_SLOTS = 4;                      // tiles in flight in tensor memory (128 columns each) = softmax warpgroups
constexpr int WT_STAGES = 7;                     // operand stages in shared memory: the TMA producer runs this many tiles ahead
constexpr int WT_THREADS = (2 + 4 * WT_SLOTS) * 32;
constexpr int WT_MAT = 128 * 64;                 // one operand matrix: 128 rows x 32 bf16 (window a: rows 0.., window b: rows 64..)
constexpr int WT_SLOT_BYTES = 3 * WT_MAT;        // Q | K | V
constexpr int WT_WIN_BYTES = 49 * 64;            // one TMA box
constexpr int WT_BIAS_LD = 52;                   // floats per bias row (49 + pad: 16-byte aligned rows, conflict-free LDS.128)
constexpr int WT_BIAS_CLS = 49 * WT_BIAS_LD;
constexpr int WT_SMEM = WT_STAGES * WT_SLOT_BYTES + 4 * WT_BIAS_CLS * 4 + (2 * WT_STAGES + 4 * WT_SLOTS) * 8 + 16 + 1024;
static_assert(WT_SMEM <= 227 * 1024, "shared memory budget");

struct WinTcParams {
  bf16* out;            // [B*H*W, C] natural token order
  const float* bias;    // [n_cls][heads][49][52] = (rel-pos bias + shift mask) * log2e
  int H, W, C, heads, shift, nWh, nWw, n_cls;
  int n_windows, n_pairs, n_groups;
  int nW_shift, nWw_shift;   // log2 of windows per image / per window row when both are powers of two (Swin at 224), else -1
  float scale_log2e;
};

__global__ void __launch_bounds__(WT_THREADS, 1)
window_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const WinTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* opnd = smem;
  float* bias_s = reinterpret_cast<float*>(smem + WT_STAGES * WT_SLOT_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 4 * WT_BIAS_CLS);
  uint64_t* full = bars;                                   // [STAGES] TMA -> MMA
  uint64_t* empty = bars + WT_STAGES;                      // [STAGES] P.V retired -> TMA (the stage's operands are free)
  uint64_t* s_full = bars + 2 * WT_STAGES;                 // [SLOTS] Q.K^T retired -> softmax warpgroup
  uint64_t* p_full = s_full + WT_SLOTS;                    // [SLOTS] P written (4 warp arrivals) -> MMA
  uint64_t* o_full = p_full + WT_SLOTS;                    // [SLOTS] P.V retired -> epilogue
  uint64_t* o_empty = o_full + WT_SLOTS;                   // [SLOTS] O read (4 warp arrivals) -> MMA may overwrite the slot's columns
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + WT_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads, group = blockIdx.x / p.heads;
  const int my_tiles = (p.n_pairs - group + p.n_groups - 1) / p.n_groups;

  // operand slots start zeroed: rows 49..63 of each window half are never written by TMA and the V rows among them are
  // multiplied (by P = 0) in the second MMA
  for (int i = threadIdx.x; i < WT_STAGES * WT_SLOT_BYTES / 16; i += WT_THREADS) reinterpret_cast<uint4*>(opnd)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    // a stage is free once P.V has retired (1 commit arrival) AND the four epilogue warps have flushed the output tile they
    // parked in its Q rows (4 arrivals)
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 5); }
    for (int s = 0; s < WT_SLOTS; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  if (warp >= 2) {   // this head's bias table (a parameter-derived constant: read before the PDL wait)
    const int nvec = p.n_cls * WT_BIAS_CLS / 4;
    for (int i = threadIdx.x - 64; i < nvec; i += WT_THREADS - 64) {
      const int cls = i / (WT_BIAS_CLS / 4), within = i - cls * (WT_BIAS_CLS / 4);
      reinterpret_cast<float4*>(bias_s)[i] =
          __ldg(reinterpret_cast<const float4*>(p.bias + ((long long)cls * p.heads + head) * WT_BIAS_CLS) + within);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_grid_sync();

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------------------------------------
    for (int n = 0; n < my_tiles; ++n) {
      const int st = n % WT_STAGES, u = n / WT_STAGES;
      mbar_wait(&empty[st], (u & 1) ^ 1);
      const int w0 = 2 * (group + n * p.n_groups);
      const int nvalid = p.n_windows - w0 >= 2 ? 2 : 1;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[st], (uint32_t)(nvalid * 3 * WT_WIN_BYTES));
        uint8_t* dst = opnd + st * WT_SLOT_BYTES;
        for (int hf = 0; hf < nvalid; ++hf)
#pragma unroll
          for (int m = 0; m < 3; ++m)
            tma_load_2d(dst + m * WT_MAT + hf * 4096, &tmap_qkv, &full[st], m * p.C + head * 32, (w0 + hf) * 49);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer: serves whichever slot is ready ---------------------------------------
    const uint32_t id_s = umma_idesc_bf16_ex(128, 128, 0), id_o = umma_idesc_bf16_ex(128, 64, 1);
    uint32_t s_cnt[WT_SLOTS], pv_cnt[WT_SLOTS];
#pragma unroll
    for (int s = 0; s < WT_SLOTS; ++s) s_cnt[s] = pv_cnt[s] = 0;
    int remaining = 2 * my_tiles;
    long long t_last = clock64();
    while (remaining > 0) {
      bool progressed = false;
#pragma unroll
      for (int slot = 0; slot < WT_SLOTS; ++slot) {
        const uint32_t n_slot = my_tiles > slot ? (uint32_t)((my_tiles - slot + WT_SLOTS - 1) / WT_SLOTS) : 0u;
        const uint32_t tm = tmem_base + slot * 128;
        if (pv_cnt[slot] < s_cnt[slot]) {
          const uint32_t u = pv_cnt[slot];
          const uint32_t n = slot + u * WT_SLOTS, st = n % WT_STAGES;          // tile -> operand stage
          const uint32_t sm = base + st * WT_SLOT_BYTES;
          if (__any_sync(0xffffffffu, mbar_test(&p_full[slot], u & 1))) {
            tc_fence_after();
            if (elect_one()) {
              // O[128 x 64] (columns 64..127) = P[128 x 64 local keys] (TMEM columns 0..31) . [V_a | V_b]
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ts_masked(tm + 64, tm + 8 * k, umma_desc(sm + 2 * WT_MAT + 1024 * k, 4096, 512, UMMA_SW64), id_o, k, 0, 0, 0, 0);
              umma_commit(&o_full[slot]);
              umma_commit(&empty[st]);
            }
            __syncwarp();
            ++pv_cnt[slot]; --remaining; progressed = true;
          }
        } else if (s_cnt[slot] < n_slot) {
          const uint32_t u = s_cnt[slot];
          const uint32_t n = slot + u * WT_SLOTS, st = n % WT_STAGES;
          const uint32_t sm = base + st * WT_SLOT_BYTES;
          if (__any_sync(0xffffffffu, mbar_test(&o_empty[slot], (u & 1) ^ 1) && mbar_test(&full[st], (n / WT_STAGES) & 1))) {
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_bf16_masked(tm, umma_desc(sm + 32 * k, 16, 512, UMMA_SW64), umma_desc(sm + WT_MAT + 32 * k, 16, 512, UMMA_SW64), id_s, k,
                                 0, 0, 0, 0);
              umma_commit(&s_full[slot]);
            }
            __syncwarp();
            ++s_cnt[slot]; --remaining; progressed = true;
          }
        }
      }
      if (progressed) t_last = clock64();
      else if (clock64() - t_last > AT_WATCHDOG) __trap();
    }
  } else {
    // ------------------------------- softmax + epilogue: one score row per thread -------------------------------------
    const int slot = (warp - 2) >> 2;
    const int quarter = warp & 3;                 // the TMEM lanes this warp may touch: [32 quarter, +32)
    const int r = quarter * 32 + lane, half = quarter >> 1, i = r & 63;
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 128;
    const int ib = i < 49 ? i : 48;
    const int ri = ib / 7, ci = ib - ri * 7;
    const int nW = p.nWh * p.nWw;
    for (int n = slot, u = 0; n < my_tiles; n += WT_SLOTS, ++u) {
      const int w = 2 * (group + n * p.n_groups) + half;
      const bool valid = w < p.n_windows && i < 49;
      const int wc = w < p.n_windows ? w : p.n_windows - 1;
      int b, win, wh, ww;
      if (p.nW_shift >= 0) {          // Swin at 224: 64 / 16 / 4 / 1 windows per image — shifts instead of divisions
        b = wc >> p.nW_shift; win = wc & (nW - 1);
        wh = win >> p.nWw_shift; ww = win & (p.nWw - 1);
      } else {
        b = wc / nW; win = wc - b * nW;
        wh = win / p.nWw; ww = win - wh * p.nWw;
      }
      // window class of the shift mask (vfe.py:321-339): windows of the last window row / column straddle the roll seam
      const int cls = p.shift > 0 ? ((wh == p.nWh - 1 ? 2 : 0) | (ww == p.nWw - 1 ? 1 : 0)) : 0;
      int hh = wh * 7 + ri + p.shift, xx = ww * 7 + ci + p.shift;     // torch.roll(-shift) then partition == read at +shift
      if (hh >= p.H) hh -= p.H;
      if (xx >= p.W) xx -= p.W;
      const int tok = valid ? (b * p.H + hh) * p.W + xx : -1;     // output row in natural token order (n_windows * 49 < 2^31)
      const float* brow = bias_s + (cls * 49 + ib) * WT_BIAS_LD;

      mbar_wait(&s_full[slot], u & 1);
      tc_fence_after();
      uint32_t a0[32], a1[16], a2;
      tmem_ld_32x32(tl + half * 64, a0);
      tmem_ld_32x16(tl + half * 64 + 32, a1);
      tmem_ld_32x1(tl + half * 64 + 48, a2);
      tmem_ld_wait();
      float s[49];
#pragma unroll
      for (int q = 0; q < 12; ++q) {
        const float4 bv = *reinterpret_cast<const float4*>(brow + 4 * q);
        const uint32_t* src = q < 8 ? &a0[4 * q] : &a1[4 * (q - 8)];
        s[4 * q + 0] = fmaf(__uint_as_float(src[0]), p.scale_log2e, bv.x);
        s[4 * q + 1] = fmaf(__uint_as_float(src[1]), p.scale_log2e, bv.y);
        s[4 * q + 2] = fmaf(__uint_as_float(src[2]), p.scale_log2e, bv.z);
        s[4 * q + 3] = fmaf(__uint_as_float(src[3]), p.scale_log2e, bv.w);
      }
      s[48] = fmaf(__uint_as_float(a2), p.scale_log2e, brow[48]);
      float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
      for (int j = 4; j < 48; j += 4) {
        m0 = fmaxf(m0, s[j]); m1 = fmaxf(m1, s[j + 1]); m2 = fmaxf(m2, s[j + 2]); m3 = fmaxf(m3, s[j + 3]);
      }
      const float mx = fmaxf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), s[48]);
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 48; j += 2) {
        const float e0 = at_ex2(s[j] - mx), e1 = at_ex2(s[j + 1] - mx);
        sum0 += e0; sum1 += e1;
        pk[j >> 1] = pack_bf16x2(e0, e1);
      }
      {
        const float e = at_ex2(s[48] - mx);
        sum0 += e;
        pk[24] = pack_bf16x2(e, 0.f);
      }
#pragma unroll
      for (int j = 25; j < 32; ++j) pk[j] = 0u;       // local keys 50..63: P = 0
      tmem_st_32x32(tl, pk);                          // P over the dead score columns 0..31 of this thread's own lane
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[slot]);
      const float inv = 1.0f / (sum0 + sum1);

      mbar_wait(&o_full[slot], u & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32(tl + 64 + half * 32, o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[slot]);
      // window_reverse + roll back == scatter through the index map.  The warp's 32 x 32 output tile is parked in the Q rows of
      // the tile's operand stage (dead since Q.K^T retired; same 64-byte rows, same swizzle) and leaves as four stores in
      // which every group of 4 lanes writes one token's 64 contiguous bytes
      {
        const int stg_n = n % WT_STAGES;
        uint8_t* q_rows = opnd + stg_n * WT_SLOT_BYTES + (quarter * 32) * 64;
        const uint32_t swz = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(q_rows + lane * 64 + (((uint32_t)q ^ swz) << 4)) =
              make_uint4(pack_bf16x2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv),
                         pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv),
                         pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv),
                         pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv));
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + (lane >> 2), cc = lane & 3;
          const int tk = __shfl_sync(0xffffffffu, tok, rr);
          const uint4 v = *reinterpret_cast<const uint4*>(q_rows + rr * 64 + (((uint32_t)cc ^ (((uint32_t)rr >> 1) & 3u)) << 4));
          if (tk >= 0) *reinterpret_cast<uint4*>(p.out + (long long)tk * p.C + head * 32 + cc * 8) = v;
        }
        __syncwarp();
        // rows 49..63 of each half must read as zero again when the stage is reused as V? no: this is the Q matrix, whose pad
        // rows only feed unused score rows — nothing to restore.  Release the stage.
        if (lane == 0) mbar_arrive(&empty[stg_n]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// BERT joint attention (head_dim 64)
// =====================================================================================================================
constexpr int JT_SLOTS = 2;
constexpr int JT_SOFTMAX_WARP0 = 3;              // warp 0: Q / K producer, warp 1: MMA issuer, warp 2: V producer + key masks
constexpr int JT_WG_WARPS = 8;                   // softmax warps per slot: TWO threads per score row (each takes half of the keys)
constexpr int JT_THREADS = (JT_SOFTMAX_WARP0 + JT_WG_WARPS * JT_SLOTS) * 32;
constexpr int JT_Q_BYTES = 128 * 128;

template <int SP, int NS> struct JtPlan {
  static constexpr int KV_BYTES = SP * 128;
  static constexpr int QK_BYTES = JT_Q_BYTES + NS * KV_BYTES;     // released as soon as Q.K^T has retired
  static constexpr int V_BYTES = NS * KV_BYTES;                   // released when P.V has retired
  static constexpr int MASK_FLOATS = JT_SLOTS * NS * SP;
  static constexpr int STAGING_BYTES = JT_WG_WARPS * JT_SLOTS * 2048;   // one [32 rows x 64 B] output tile per softmax warp
  static constexpr int NCS = SP / 16;                             // 16-key chunks of a score row
  static constexpr int XCH_SLOT = 256 + 128 * (SP / 16);          // per slot: partial row maxima [128][2], per-chunk row sums [128][NCS]
  static constexpr int XCH_FLOATS = JT_SLOTS * XCH_SLOT;
  static constexpr int KSPLIT = SP == 144 ? 80 : 48;              // keys [0, KSPLIT) -> thread half 0, [KSPLIT, SP) -> half 1
  static constexpr int SMEM = JT_SLOTS * (QK_BYTES + V_BYTES) + STAGING_BYTES + (MASK_FLOATS + XCH_FLOATS) * 4 + 9 * JT_SLOTS * 8 + 16 + 1024;
  static_assert(KSPLIT % 16 == 0 && (SP - KSPLIT) % 16 == 0, "each half is a whole number of 16-column TMEM loads");
  static_assert(KV_BYTES % 1024 == 0 && SP % 16 == 0 && SP <= 144, "K / V buffers are whole SWIZZLE_128B atoms; S columns + O fit 256");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct JointTcParams {
  bf16* out;              // [R, heads*64]
  const float* kmask;     // [B, S] additive (0 / -10000)
  int R, S, B, heads, seq2seq, obj_end, n_items;
  float scale;
  unsigned long long* trace;   // debug: clock64 stamps of CTA 0 (tools/attn_trace.py); nullptr in production
};
static unsigned long long* g_attn_trace = nullptr;
// slot t of event e of tile n (first 8 tiles of CTA 0): trace[16 + n * 16 + e]; trace[0] = kernel start
#define JT_STAMP(n, e) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && (n) < 8) p.trace[16 + (n) * 16 + (e)] = (unsigned long long)clock64(); } while (0)

// bits [lo, hi) of a 128-bit lane set, word w
__device__ __forceinline__ uint32_t jt_range_word(int lo, int hi, int w) {
  const int a = lo - 32 * w, b = hi - 32 * w;
  const int l = a < 0 ? 0 : a, h = b > 32 ? 32 : b;
  if (h <= l) return 0u;
  const uint32_t upto_h = h >= 32 ? 0xffffffffu : ((1u << h) - 1u);
  const uint32_t upto_l = l >= 32 ? 0xffffffffu : ((1u << l) - 1u);
  return upto_h & ~upto_l;
}

// one 32- or 16-column chunk of a score row: v_j = acc_j * scale + mask_j (the reference's units), optional seq2seq blocking
template <int W>
__device__ __forceinline__ void jt_scores(const uint32_t (&a)[W], float (&v)[W], const float* mrow, float scale, bool seq2seq, int j0, int i,
                                          int obj_end) {
  const float blocked = -10000.0f;
#pragma unroll
  for (int q = 0; q < W / 4; ++q) {
    const float4 mv = *reinterpret_cast<const float4*>(mrow + j0 + 4 * q);
    v[4 * q + 0] = fmaf(__uint_as_float(a[4 * q + 0]), scale, mv.x);
    v[4 * q + 1] = fmaf(__uint_as_float(a[4 * q + 1]), scale, mv.y);
    v[4 * q + 2] = fmaf(__uint_as_float(a[4 * q + 2]), scale, mv.z);
    v[4 * q + 3] = fmaf(__uint_as_float(a[4 * q + 3]), scale, mv.w);
  }
  if (seq2seq) {   // model.py:118-123: text rows see the image block and the text up to themselves
#pragma unroll
    for (int c = 0; c < W; ++c)
      if (j0 + c > i && j0 + c > obj_end) v[c] += blocked;
  }
}

template <int SP, int NS>
__global__ void __launch_bounds__(JT_THREADS, 1)
joint_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_out, const JointTcParams p) {
  using P = JtPlan<SP, NS>;
  constexpr int KV = P::KV_BYTES, QKB = P::QK_BYTES, VB = P::V_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smem_v = smem + JT_SLOTS * QKB;
  uint8_t* staging = smem_v + JT_SLOTS * VB;       // 1024-byte aligned: QKB and VB are multiples of 1024
  float* mask_s = reinterpret_cast<float*>(staging + P::STAGING_BYTES);
  float* xch = mask_s + P::MASK_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + P::XCH_FLOATS);
  uint64_t* qk_full = bars;                      // TMA (Q, K) -> MMA
  uint64_t* qk_empty = bars + JT_SLOTS;          // Q.K^T retired -> Q / K producer
  uint64_t* v_full = bars + 2 * JT_SLOTS;        // TMA (V) -> MMA
  uint64_t* v_empty = bars + 3 * JT_SLOTS;       // P.V retired -> V producer
  uint64_t* s_full = bars + 4 * JT_SLOTS;        // Q.K^T retired -> softmax warpgroup
  uint64_t* p_full = bars + 5 * JT_SLOTS;        // P written (4 warp arrivals) -> MMA
  uint64_t* o_full = bars + 6 * JT_SLOTS;        // P.V retired -> epilogue
  uint64_t* o_empty = bars + 7 * JT_SLOTS;       // O read (4 warp arrivals) -> MMA may overwrite the slot's TMEM columns
  uint64_t* m_full = bars + 8 * JT_SLOTS;        // key masks of the slot's tile staged by the V-producer warp -> softmax warpgroup
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9 * JT_SLOTS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int C = p.heads * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < JT_SLOTS; ++s) {
      mbar_init(&qk_full[s], 1); mbar_init(&qk_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], JT_WG_WARPS); mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], JT_WG_WARPS);
      mbar_init(&m_full[s], 32);                  // every lane of the V-producer warp: cp.async.mbarrier.arrive.noinc
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  // key-mask rows start as 0 for the keys of a sequence and "minus infinity" past S; the per-tile copies only touch j < S
  for (int i = threadIdx.x; i < P::MASK_FLOATS; i += JT_THREADS) mask_s[i] = (i % SP) < p.S ? 0.f : AT_NEG_BIG;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_grid_sync();
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[0] = (unsigned long long)clock64();

  // item n of this CTA -> (row tile, head): heads vary fastest, so the CTAs of a wave share the row tile's K / V rows in L2
  auto item_of = [&](int n, int& row0, int& head, int& b0, int& ns) {
    const int it = (int)blockIdx.x + n * (int)gridDim.x;
    head = it % p.heads;
    row0 = (it / p.heads) * 128;
    b0 = row0 / p.S;
    const int last = row0 + 127 < p.R ? row0 + 127 : p.R - 1;
    ns = last / p.S - b0 + 1;
  };

  if (warp == 0) {
    // ------------------------------- TMA producer: Q and K (their buffers are recycled as soon as Q.K^T has retired, so
    // the next tile of the slot is resident long before the softmax of the current one ends) --------------------------
    for (int n = 0; n < my_tiles; ++n) {
      const int slot = n % JT_SLOTS, u = n / JT_SLOTS;
      int row0, head, b0, ns;
      item_of(n, row0, head, b0, ns);
      mbar_wait(&qk_empty[slot], (u & 1) ^ 1);
      JT_STAMP(n, 0);
      if (elect_one()) {
        uint8_t* dst = smem + slot * QKB;
        mbar_arrive_expect_tx(&qk_full[slot], (uint32_t)(JT_Q_BYTES + ns * KV));
        tma_load_2d(dst, &tmap_q, &qk_full[slot], head * 64, row0);                 // rows past R arrive as zeros
        // SP rows from the sample's first row: the rows past S belong to the next sample (finite, masked) or are zero fill
        for (int k = 0; k < ns; ++k) tma_load_2d(dst + JT_Q_BYTES + k * KV, &tmap_kv, &qk_full[slot], C + head * 64, (b0 + k) * p.S);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ------------------------------- TMA producer: V; it also stages the additive key masks of the tile's samples (log2
    // units, keys past S = "minus infinity") one tile ahead of the softmax warps.  v_empty of the slot implies that P.V of
    // its previous tile has retired, i.e. every reader of the previous masks is done. ------------------------------------
    const bool s2s_ = p.seq2seq != 0;
    for (int n = 0; n < my_tiles; ++n) {
      const int slot = n % JT_SLOTS, u = n / JT_SLOTS;
      int row0, head, b0, ns;
      item_of(n, row0, head, b0, ns);
      mbar_wait(&v_empty[slot], (u & 1) ^ 1);
      JT_STAMP(n, 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[slot], (uint32_t)(ns * KV));
        for (int k = 0; k < ns; ++k) tma_load_2d(smem_v + slot * VB + k * KV, &tmap_kv, &v_full[slot], 2 * C + head * 64, (b0 + k) * p.S);
      }
      __syncwarp();
      // additive key masks (0 / -10000, model.py:126) of the tile's samples: asynchronous 4-byte global -> shared copies (no
      // registers, all in flight at once); their completion arrives on m_full.  seq2seq: the rows stay 0 (mask is positional)
      float* msk = mask_s + slot * NS * SP;
      if (!s2s_) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          const int bb = b0 + k;
          if (bb < p.B)
            for (int j = lane; j < p.S; j += 32)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(msk + k * SP + j)), "l"(p.kmask + (long long)bb * p.S + j) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&m_full[slot])) : "memory");
      JT_STAMP(n, 4);
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------------------------------------
    const uint32_t id_s = umma_idesc_bf16_ex(128, SP, 0), id_o = umma_idesc_bf16_ex(128, 64, 1);
    uint32_t s_cnt[JT_SLOTS], pv_cnt[JT_SLOTS];
#pragma unroll
    for (int s = 0; s < JT_SLOTS; ++s) s_cnt[s] = pv_cnt[s] = 0;
    int remaining = 2 * my_tiles;
    long long t_last = clock64();
    while (remaining > 0) {
      bool progressed = false;
#pragma unroll
      for (int slot = 0; slot < JT_SLOTS; ++slot) {
        const uint32_t n_slot = my_tiles > slot ? (uint32_t)((my_tiles - slot + JT_SLOTS - 1) / JT_SLOTS) : 0u;
        const uint32_t tm = tmem_base + slot * 256;
        const uint32_t sm_qk = base + slot * QKB, sm_v = base + JT_SLOTS * QKB + slot * VB;
        const bool do_pv = pv_cnt[slot] < s_cnt[slot];
        if (!do_pv && s_cnt[slot] >= n_slot) continue;
        const uint32_t u = do_pv ? pv_cnt[slot] : s_cnt[slot];
        // Q.K^T of the slot's NEXT tile only overwrites the score columns, whose last reader is the P.V just issued in front of
        // it (same thread: the tensor pipe keeps issue order) — it does not wait for the epilogue.  P.V writes the O columns
        // and does: o_empty.
        const bool ready = do_pv ? (mbar_test(&p_full[slot], u & 1) && mbar_test(&v_full[slot], u & 1) && mbar_test(&o_empty[slot], (u & 1) ^ 1))
                                 : mbar_test(&qk_full[slot], u & 1);
        if (!__any_sync(0xffffffffu, ready)) {
          JT_STAMP(slot + (int)u * JT_SLOTS, do_pv ? 12 : 11);     // last poll that found the step not ready
          continue;
        }
        tc_fence_after();
        int row0, head, b0, ns;
        item_of(slot + (int)u * JT_SLOTS, row0, head, b0, ns);
        JT_STAMP(slot + (int)u * JT_SLOTS, do_pv ? 3 : 2);
        if (elect_one()) {
          for (int k = 0; k < ns; ++k) {
            // lanes of this tile that belong to sample b0 + k; every other lane is disabled for this sample's products
            const int lo = (b0 + k) * p.S - row0, hi = (b0 + k + 1) * p.S - row0;
            const uint32_t d0 = ~jt_range_word(lo, hi, 0), d1 = ~jt_range_word(lo, hi, 1), d2 = ~jt_range_word(lo, hi, 2),
                           d3 = ~jt_range_word(lo, hi, 3);
            if (!do_pv) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_bf16_masked(tm, umma_desc(sm_qk + 32 * ks, 16, 1024, UMMA_SW128),
                                 umma_desc(sm_qk + JT_Q_BYTES + k * KV + 32 * ks, 16, 1024, UMMA_SW128), id_s, ks, d0, d1, d2, d3);
            } else {
#pragma unroll
              for (int ks = 0; ks < SP / 16; ++ks)
                umma_bf16_ts_masked(tm + 192, tm + 8 * ks, umma_desc(sm_v + k * KV + 2048 * ks, 16, 1024, UMMA_SW128), id_o, ks, d0, d1, d2, d3);
            }
          }
          if (do_pv) {
            umma_commit(&o_full[slot]);
            umma_commit(&v_empty[slot]);
          } else {
            umma_commit(&s_full[slot]);
            umma_commit(&qk_empty[slot]);
          }
        }
        __syncwarp();
        if (do_pv) ++pv_cnt[slot]; else ++s_cnt[slot];
        --remaining;
        progressed = true;
      }
      if (progressed) t_last = clock64();
      else if (clock64() - t_last > AT_WATCHDOG) __trap();
    }
  } else {
    // ------------------------------- softmax + epilogue: TWO threads per score row -------------------------------------
    // Warps 3 + 8 slot .. : the two warps that own the same TMEM lane quarter split the keys (thread half 0: keys [0, KSPLIT),
    // half 1: the rest) — half the dependent chain per thread and twice the warps to hide tcgen05.ld / MUFU latency.  Partial row
    // maxima and sums are exchanged through shared memory around two named barriers of the slot's 256 threads; the second
    // one also separates every read of the score columns from the P stores that alias them.
    constexpr int KSPLIT = P::KSPLIT;
    const int wsub = (warp - JT_SOFTMAX_WARP0) % JT_WG_WARPS;
    const int slot = (warp - JT_SOFTMAX_WARP0) / JT_WG_WARPS;
    const int quarter = warp & 3;
    const int half = wsub >> 2;                      // warps 3..6 -> quarters 3,0,1,2 (half 0), warps 7..10 -> the same quarters (half 1)
    const int r = quarter * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 256;
    float* msk = mask_s + slot * NS * SP;
    constexpr int NCS = P::NCS;
    float* x_max = xch + slot * P::XCH_SLOT;         // [128][2]
    float* x_cs = x_max + 256;                       // [128][NCS]: sum of exp over each 16-key chunk of the row
    uint8_t* stg = staging + (warp - JT_SOFTMAX_WARP0) * 2048;     // this warp's output tile: 32 rows x 64 B, SWIZZLE_64B
    const uint32_t stg_row = (uint32_t)lane * 64u, stg_swz = ((uint32_t)lane >> 1) & 3u;
    const bool s2s = p.seq2seq != 0;
    const int k_lo = half ? KSPLIT : 0;              // this thread's keys [k_lo, k_lo + NK)
    constexpr int NK0 = KSPLIT, NK1 = SP - KSPLIT;
    for (int n = slot, u = 0; n < my_tiles; n += JT_SLOTS, ++u) {
      int row0, head, b0, ns;
      item_of(n, row0, head, b0, ns);
      const int g = row0 + r;
      const bool valid = g < p.R;
      const int b = valid ? g / p.S : b0;
      const int i = g - b * p.S;                         // position of this row in its sample
      const float* mrow = msk + (b - b0) * SP;
      mbar_wait(&m_full[slot], u & 1);
      if (wsub == 1) JT_STAMP(n, 5);

      mbar_wait(&s_full[slot], u & 1);
      tc_fence_after();
      if (wsub == 1) JT_STAMP(n, 6);
      // pass 1: maximum over this thread's keys, 16 columns per TMEM load
      float mx = AT_NEG_BIG;
      const int nk = half ? NK1 : NK0;
#pragma unroll 1
      for (int c0 = 0; c0 < nk; c0 += 16) {
        uint32_t a[16];
        float v[16];
        tmem_ld_32x16(tl + k_lo + c0, a);
        tmem_ld_wait();
        jt_scores<16>(a, v, mrow, p.scale, s2s, k_lo + c0, i, p.obj_end);
        float m0 = fmaxf(v[0], v[1]), m1 = fmaxf(v[2], v[3]), m2 = fmaxf(v[4], v[5]), m3 = fmaxf(v[6], v[7]);
#pragma unroll
        for (int c = 8; c < 16; c += 4) { m0 = fmaxf(m0, v[c]); m1 = fmaxf(m1, v[c + 1]); m2 = fmaxf(m2, v[c + 2]); m3 = fmaxf(m3, v[c + 3]); }
        mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      }
      x_max[r * 2 + half] = mx;
      named_bar_sync(1 + slot, JT_WG_WARPS * 32);
      mx = fmaxf(mx, x_max[r * 2 + (half ^ 1)]);
      if (wsub == 1) JT_STAMP(n, 7);
      // pass 2: exp2((v - max) * log2e), partial row sum; P as bf16 pairs, held in registers until every thread of the slot
      // has finished reading the score columns
      // The row sum is accumulated per 16-key chunk and the chunk sums are added in key order by BOTH threads of the row, so the
      // result does not depend on where the keys are split between them (nor on SP: padded chunks add exact zeros) — a
      // sequence and its [PAD]-extended copy give bit-identical rows
      const float nmx = -mx * AT_LOG2E;
      uint32_t pk[NK0 / 2];
#pragma unroll
      for (int cc = 0; cc < NK0 / 16; ++cc) {
        if (cc * 16 < nk) {       // warp-uniform
          uint32_t a[16];
          float v[16];
          tmem_ld_32x16(tl + k_lo + cc * 16, a);
          tmem_ld_wait();
          jt_scores<16>(a, v, mrow, p.scale, s2s, k_lo + cc * 16, i, p.obj_end);
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            const float e0 = at_ex2(fmaf(v[c], AT_LOG2E, nmx)), e1 = at_ex2(fmaf(v[c + 1], AT_LOG2E, nmx));
            sum0 += e0; sum1 += e1;
            pk[cc * 8 + (c >> 1)] = pack_bf16x2(e0, e1);
          }
          x_cs[r * NCS + (k_lo >> 4) + cc] = sum0 + sum1;
        }
      }
      named_bar_sync(1 + slot, JT_WG_WARPS * 32);
      // P columns: key pair (2q, 2q + 1) -> TMEM column q of the slot (aliasing the score columns, now dead)
#pragma unroll
      for (int cc = 0; cc < NK0 / 16; ++cc) {
        if (cc * 16 < nk) {
          uint32_t t8[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) t8[q] = pk[cc * 8 + q];
          tmem_st_32x8(tl + ((k_lo + cc * 16) >> 1), t8);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[slot]);
      float rsum = 0.f;
#pragma unroll
      for (int c = 0; c < NCS; ++c) rsum += x_cs[r * NCS + c];
      const float inv = 1.0f / rsum;
      if (wsub == 1) JT_STAMP(n, 8);

      mbar_wait(&o_full[slot], u & 1);
      tc_fence_after();
      if (wsub == 1) JT_STAMP(n, 9);
      uint32_t o0[32];
      tmem_ld_32x32(tl + 192 + half * 32, o0);          // this thread's 32 of the row's 64 output columns
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[slot]);
      // the warp's 32 x 32 output tile goes through shared memory (TMA swizzle pattern, conflict-free 16-byte stores) and
      // leaves as ONE tensor store of 32 rows x 64 bytes; rows past R are clipped
      if (lane == 0) bulk_wait_read<0>();             // the previous tile's store has finished reading the staging buffer
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(stg + stg_row + (((uint32_t)q ^ stg_swz) << 4)) =
            make_uint4(pack_bf16x2(__uint_as_float(o0[8 * q]) * inv, __uint_as_float(o0[8 * q + 1]) * inv),
                       pack_bf16x2(__uint_as_float(o0[8 * q + 2]) * inv, __uint_as_float(o0[8 * q + 3]) * inv),
                       pack_bf16x2(__uint_as_float(o0[8 * q + 4]) * inv, __uint_as_float(o0[8 * q + 5]) * inv),
                       pack_bf16x2(__uint_as_float(o0[8 * q + 6]) * inv, __uint_as_float(o0[8 * q + 7]) * inv));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (row0 + quarter * 32 < p.R) tma_store_2d(&tmap_out, stg, head * 64 + half * 32, row0 + quarter * 32);
        bulk_commit();
      }
      __syncwarp();
      if (wsub == 1) JT_STAMP(n, 10);
    }
  }

  if (warp >= JT_SOFTMAX_WARP0 && lane == 0) bulk_wait_all();     // this warp's output stores are globally performed
  tc_fence_before();
  __syncthreads();
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.trace[1] = (unsigned long long)clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int SP, int NS>
static int launch_joint_tc(const void* qkv, const JointTcParams& p, cudaStream_t stream) {
  using P = JtPlan<SP, NS>;
  static unsigned long long attr_devices = 0;
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(joint_attn_tc_kernel<SP, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM);
    if (e != cudaSuccess) return (int)e;
  }
  const int C = p.heads * 64;
  CUtensorMap tq, tkv;
  int rc;
  if ((rc = make_tmap(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, p.R, 3 * C, 3 * C, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, p.R, 3 * C, 3 * C, 64, SP, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) != MVLT_OK) return rc;
  CUtensorMap tout;
  if ((rc = make_tmap(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.out, p.R, C, C, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE)) != MVLT_OK) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_items < sms ? p.n_items : sms;
  cudaError_t e = launch_k(joint_attn_tc_kernel<SP, NS>, dim3(grid), dim3(JT_THREADS), (size_t)P::SMEM, stream, tq, tkv, tout, p);
  return e == cudaSuccess ? MVLT_OK : (int)e;
}

}  // namespace mvlt

using namespace mvlt;

// debug hook (not part of include/mvlt_b200.h): device buffer of >= 256 u64 stamped by CTA 0 of every later joint-attention launch
extern "C" int mvlt_debug_attn_trace(void* dev_buf) {
  g_attn_trace = reinterpret_cast<unsigned long long*>(dev_buf);
  return MVLT_OK;
}

// Swin window attention on the tensor cores.  qkv: bf16 [B*nW*49, 3C], rows WINDOW-MAJOR for this block's shift (row
// (b*nW + w)*49 + i = token i of window w of the rolled image b); out: bf16 [B*H*W, C], natural token order.
// bias_table: fp32 [n_cls][heads][49][52], (rel-pos bias + shift mask) * log2(e), n_cls = 4 when shift > 0 else 1
// (ops.window_bias_table).  Replaces vfe.py:234-251 + the roll / partition / reverse of :361-378.
extern "C" int mvlt_window_attention_tc(const void* qkv, void* out, const float* bias_table, int B, int H, int W, int C, int heads,
                                        int window, int shift, float scale, cudaStream_t stream) {
  if (!qkv || !out || !bias_table || B <= 0 || heads <= 0 || C != heads * 32) return MVLT_ERR_INVALID;
  if (window != 7 || H % window || W % window || shift < 0 || shift >= window) return MVLT_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15) || ((uintptr_t)bias_table & 15)) return MVLT_ERR_INVALID;
  int rc = mvlt_gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  static unsigned long long attr_devices = 0;
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM);
    if (e != cudaSuccess) return (int)e;
  }
  WinTcParams p;
  p.out = reinterpret_cast<bf16*>(out); p.bias = bias_table;
  p.H = H; p.W = W; p.C = C; p.heads = heads; p.shift = shift; p.nWh = H / window; p.nWw = W / window;
  p.n_cls = shift > 0 ? 4 : 1;
  const long long n_windows = (long long)B * p.nWh * p.nWw;
  if (n_windows * 49 > 0x7fffffffLL) return MVLT_ERR_UNSUPPORTED;
  p.n_windows = (int)n_windows;
  p.n_pairs = (p.n_windows + 1) / 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int groups = sms / heads;
  if (groups < 1) groups = 1;
  if (groups > p.n_pairs) groups = p.n_pairs;
  p.n_groups = groups;
  auto log2_exact = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  p.nW_shift = log2_exact(p.nWh * p.nWw);
  p.nWw_shift = log2_exact(p.nWw);
  if (p.nW_shift < 0 || p.nWw_shift < 0) p.nW_shift = p.nWw_shift = -1;
  p.scale_log2e = scale * AT_LOG2E;
  CUtensorMap tm;
  if ((rc = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, n_windows * 49, 3LL * C, 3LL * C, 32, 49, CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) != MVLT_OK) return rc;
  cudaError_t e = launch_k(window_attn_tc_kernel, dim3(groups * heads), dim3(WT_THREADS), (size_t)WT_SMEM, stream, tm, p);
  return e == cudaSuccess ? MVLT_OK : (int)e;
}

// BERT joint attention on the tensor cores.  qkv: bf16 [B*S, 3*heads*64] (q | k | v), out: bf16 [B*S, heads*64]; kmask fp32 [B, S]
// additive (ignored when seq2seq).  Sequence lengths 64..96 and 128..144 (the MVLT joint sequences: 74 / 81 / 131); other
// lengths return MVLT_ERR_UNSUPPORTED (callers fall back to mvlt_joint_attention).  Replaces HF modeling_bert.py:115-140.
extern "C" int mvlt_joint_attention_tc(const void* qkv, void* out, const float* kmask, int B, int S, int heads, int head_dim,
                                       int seq2seq, int obj_end, float scale, cudaStream_t stream) {
  if (!qkv || !out || !kmask || B <= 0 || S <= 0 || heads <= 0) return MVLT_ERR_INVALID;
  if (head_dim != 64) return MVLT_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return MVLT_ERR_INVALID;
  const long long R = (long long)B * S;
  if (R > 0x7fffffffLL - 256) return MVLT_ERR_UNSUPPORTED;
  int rc = mvlt_gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  JointTcParams p;
  p.out = reinterpret_cast<bf16*>(out); p.kmask = kmask; p.R = (int)R; p.S = S; p.B = B; p.heads = heads; p.seq2seq = seq2seq;
  p.obj_end = obj_end; p.scale = scale;
  p.trace = g_attn_trace;
  p.n_items = (int)((R + 127) / 128) * heads;
  if (S >= 128 && S <= 144) return launch_joint_tc<144, 2>(qkv, p, stream);
  if (S >= 64 && S <= 96) return launch_joint_tc<96, 3>(qkv, p, stream);
  return MVLT_ERR_UNSUPPORTED;
}
