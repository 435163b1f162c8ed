// First half of a Swin block up to the attention input as ONE tcgen05 kernel on CTA pairs (cta_group::2):
//
//     qkv[window-major rows, 3C] (bf16) = LayerNorm1(x)[rolled + window-partitioned rows, C] . Wqkv^T + b_qkv
//
// vfe.py:356 (norm1), :361 (torch.roll), :363-364 (window_partition), :231 (qkv Linear).  A cluster of two CTAs owns 256
// OUTPUT rows, which are consecutive rows of the window-major order the tcgen05 window-attention kernel reads; the LayerNorm
// prologue GATHERS the matching token rows of the natural-order fp32 residual stream (one warp per row, 128-bit loads — a
// row is 4C contiguous bytes wherever it lives), normalises them (fp32 statistics, two passes from registers) and writes
// them as the bf16 K-major SWIZZLE_128B A operand, which then STAYS in shared memory while the 3C output columns are walked in
// 256-wide chunks: each CTA TMA-loads half of every [256, 64] weight tile, one thread of the leader issues M = 256 MMAs into
// one of two TMEM accumulator buffers, and the sixteen compute warps drain the other (tcgen05.ld -> + bias -> bf16 ->
// swizzled staging -> TMA store).  Against LayerNorm kernel + GEMM this removes the bf16 LayerNorm round trip (4C bytes per
// row), one launch per block and the GEMM's re-read of A for every column tile.
#include <cuda.h>

#include "common.cuh"
#include "cg2.cuh"
#include "tmap.cuh"

extern "C" int mvlt_gemm_tc_init(void);

namespace mvlt {

constexpr int LQ_THREADS = 18 * 32;        // warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-17 compute
constexpr int LQ_NCW = 16;
constexpr int LQ_SLOT = 128 * 128;         // one CTA's half of a [256, 64] weight tile

template <int C> struct LnQkvPlan {
  static_assert(C == 384 || C == 192 || C == 96, "Swin-S stage 0 / 1 / 2 widths");
  static constexpr int KB1 = (C + 63) / 64;   // C = 96: the second k-block is half empty (zero A columns, zero-filled weight columns)
  static constexpr int A1_BYTES = KB1 * 16384;
  static constexpr int STAGING_BYTES = LQ_NCW * 2048;
  static constexpr int PAR_BYTES = 2 * C * 4;
  static constexpr int NSLOT_RAW = (227 * 1024 - 1024 - 512 - A1_BYTES - STAGING_BYTES - PAR_BYTES) / LQ_SLOT;
  static constexpr int NSLOT = NSLOT_RAW > 6 ? 6 : NSLOT_RAW;
  static constexpr int SMEM = A1_BYTES + NSLOT * LQ_SLOT + STAGING_BYTES + PAR_BYTES + 512 + 1024;
  static_assert(NSLOT >= 4 && SMEM <= 227 * 1024, "shared memory budget");
};

struct LnQkvParams {
  const float* x;        // [B*H*W, C] fp32, natural token order (row stride ldx)
  long long ldx;
  const float* gamma;
  const float* beta;
  const float* bias;     // [N]
  float eps;
  int M, N;              // rows (= B*H*W), output columns (= 3C)
  int H, W, shift, nWw, nW;
  int nW_shift, nWw_shift;   // log2 of the window counts when both are powers of two (Swin at 224), else -1
  unsigned long long* trace;   // debug: clock64 stamps of CTA 0 (tools/lnqkv_trace.py); nullptr in production
};
static unsigned long long* g_lnqkv_trace = nullptr;
#define LQ_STAMP(idx) do { if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[(idx)] = (unsigned long long)clock64(); } while (0)

// window-major row R -> natural token row (inverse of the WinMap of rowwise.cu), 7x7 windows.  Divisions by the constants 49
// and 7 compile to multiplies; the window counts are powers of two for Swin at 224 (the first version's runtime divisions
// cost 300 instructions per row: 13k of the kernel's 32k cycles, profiles/r02_lnqkv_trace.log)
constexpr int LQ_WS = 7;
__device__ __forceinline__ long long lq_token_of(const LnQkvParams& p, int R) {
  const int wg = R / (LQ_WS * LQ_WS), i = R - wg * (LQ_WS * LQ_WS);
  int b, w, wh, ww;
  if (p.nW_shift >= 0) {
    b = wg >> p.nW_shift; w = wg & (p.nW - 1);
    wh = w >> p.nWw_shift; ww = w & (p.nWw - 1);
  } else {
    b = wg / p.nW; w = wg - b * p.nW;
    wh = w / p.nWw; ww = w - wh * p.nWw;
  }
  const int r = i / LQ_WS, c = i - r * LQ_WS;
  int h = wh * LQ_WS + r + p.shift, xx = ww * LQ_WS + c + p.shift;    // the rolled image reads the token at +shift
  if (h >= p.H) h -= p.H;
  if (xx >= p.W) xx -= p.W;
  return ((long long)b * p.H + h) * p.W + xx;
}

template <int C>
__global__ void __launch_bounds__(LQ_THREADS, 1)
swin_ln_qkv_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out, const LnQkvParams p) {
  using P = LnQkvPlan<C>;
  constexpr int KB1 = P::KB1, NSLOT = P::NSLOT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a1 = smem;
  uint8_t* ring = a1 + P::A1_BYTES;
  uint8_t* staging = ring + NSLOT * LQ_SLOT;
  float* par = reinterpret_cast<float*>(staging + P::STAGING_BYTES);     // gamma[C] | beta[C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(par + 2 * C);
  uint64_t* w_full = bars;               // [NSLOT] TMA -> MMA            (leader counts both CTAs' bytes)
  uint64_t* w_empty = w_full + NSLOT;    // [NSLOT] MMA -> TMA            (multicast commit)
  uint64_t* a1_full = w_empty + NSLOT;   // LayerNorm rows written        (leader, 2 x 16 warp arrivals)
  uint64_t* acc_full = a1_full + 1;      // [2] chunk accumulated         (multicast commit)
  uint64_t* acc_empty = acc_full + 2;    // [2] chunk drained             (leader, 2 x 16 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (blockIdx.x / 2) * 256 + (int)rank * 128;      // first (window-major) output row of this CTA
  const int n_chunks = (p.N + 255) / 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a1_full, 2 * LQ_NCW);
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2 * LQ_NCW); }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_ptr, 512);
    tmem_relinquish_cg2();
  }
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < C; i += LQ_NCW * 32) { par[i] = __ldg(p.gamma + i); par[C + i] = __ldg(p.beta + i); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------- TMA producer: weights only (static: no dependency wait) -------------------------
    uint32_t cnt = 0;
    for (int nc = 0; nc < n_chunks; ++nc) {
      const int nw = p.N - nc * 256 >= 256 ? 256 : p.N - nc * 256;     // chunk width; each CTA supplies nw / 2 weight rows
      for (int kb = 0; kb < KB1; ++kb, ++cnt) {
        const uint32_t s = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
        mbar_wait(&w_empty[s], ph ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&w_full[s], 2 * LQ_SLOT);        // rows past N arrive as zeros: the box is the byte count
          tma_load_cg2(ring + s * LQ_SLOT, &tmap_w, &w_full[s], kb * 64, nc * 256 + (int)rank * (nw / 2));
        }
        __syncwarp();
      }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA) ----------------------------------------------------------
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (rank == 0) {
      constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);          // SBO 1024 B | version 1 | SWIZZLE_128B
      const uint32_t lo_a1 = (smem_u32(a1) >> 4) | (1u << 16);
      const uint32_t lo_w = (smem_u32(ring) >> 4) | (1u << 16);
      LQ_STAMP(0);
      mbar_wait(a1_full, 0);
      tc_fence_after();
      LQ_STAMP(1);
      uint32_t cnt = 0;
      for (int nc = 0; nc < n_chunks; ++nc) {
        const int b = nc & 1, u = nc >> 1;
        const int nw = p.N - nc * 256 >= 256 ? 256 : p.N - nc * 256;
        const uint32_t idesc = umma_idesc_bf16(256, (uint32_t)nw);
        mbar_wait(&acc_empty[b], (u & 1) ^ 1);                   // the compute warps have drained this buffer (chunk nc - 2)
        tc_fence_after();
        LQ_STAMP(16 + 2 * nc);
        const uint32_t d = tmem_base + b * 256;
        for (int kb = 0; kb < KB1; ++kb, ++cnt) {
          const uint32_t s = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
          mbar_wait(&w_full[s], ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_cg2(d, ((uint64_t)DESC_HI << 32) | (lo_a1 + kb * (16384 >> 4) + 2 * k),
                            ((uint64_t)DESC_HI << 32) | (lo_w + s * (LQ_SLOT >> 4) + 2 * k), idesc, (kb | k) != 0);
            umma_commit_cg2(&w_empty[s]);
            if (kb == KB1 - 1) umma_commit_cg2(&acc_full[b]);
          }
          __syncwarp();
        }
        LQ_STAMP(16 + 2 * nc + 1);
      }
    }
  } else {
    // ------------------------------- compute warps: LayerNorm prologue (warp per row), then the epilogue --------------
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes [32 quarter, +32)
    const int part = ew >> 2;      // this warp's 32-column chunks of a 256-column buffer: part, part + 4
    pdl_grid_sync();               // x is written by the previous kernel
    if (ew == 0) LQ_STAMP(2);
    {
      constexpr int NCH = (C / 4 + 31) / 32;  // float4 per lane per row (the last one partial when C % 128 != 0)
      constexpr int ROWS = 128 / LQ_NCW;      // 8 rows per warp
      constexpr int RU = 4;
      constexpr float inv_c = 1.0f / (float)C;
#pragma unroll 1
      for (int rr = 0; rr < ROWS; rr += RU) {
        float4 v[RU][NCH];
        // lane u resolves output row u of this batch to its token; broadcast below
        const int Rl = row0 + ew * ROWS + rr + (lane & (RU - 1));
        const long long tok_l = Rl < p.M ? lq_token_of(p, Rl) : -1;
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const long long tok = __shfl_sync(0xffffffffu, tok_l, u);
          const bool ok = tok >= 0;
          const float* src = p.x + (ok ? tok : 0) * p.ldx;
#pragma unroll
          for (int i = 0; i < NCH; ++i)
            v[u][i] = (ok && (lane + 32 * i) * 4 < C) ? load4(src + (lane + 32 * i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mean[RU], rstd[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < NCH; ++i) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
          mean[u] = s;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < RU; ++u) mean[u] += __shfl_xor_sync(0xffffffffu, mean[u], o);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          mean[u] *= inv_c;
          float q = 0.f;
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            if ((lane + 32 * i) * 4 < C) {
              const float a = v[u][i].x - mean[u], b = v[u][i].y - mean[u], cc = v[u][i].z - mean[u], d = v[u][i].w - mean[u];
              q += (a * a + b * b) + (cc * cc + d * d);
            }
          }
          rstd[u] = q;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < RU; ++u) rstd[u] += __shfl_xor_sync(0xffffffffu, rstd[u], o);
#pragma unroll
        for (int u = 0; u < RU; ++u) rstd[u] = 1.0f / sqrtf(rstd[u] * inv_c + p.eps);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c = (lane + 32 * i) * 4;
          if (c >= KB1 * 64) continue;
          const bool pad = c >= C;                 // columns C .. 64 KB1 of the last k-block (C = 96): zeros, the MMAs read them
          const float4 g = pad ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(par + c);
          const float4 b = pad ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(par + C + c);
          const int kb = c >> 6, within = c & 63;
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int r = ew * ROWS + rr + u;
            const float o0 = (v[u][i].x - mean[u]) * rstd[u] * g.x + b.x;
            const float o1 = (v[u][i].y - mean[u]) * rstd[u] * g.y + b.y;
            const float o2 = (v[u][i].z - mean[u]) * rstd[u] * g.z + b.z;
            const float o3 = (v[u][i].w - mean[u]) * rstd[u] * g.w + b.w;
            const uint32_t off = (uint32_t)kb * 16384u + (uint32_t)r * 128u +
                                 ((((uint32_t)(within >> 3)) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)within & 4u) << 1);
            *reinterpret_cast<uint2*>(a1 + off) = pad ? make_uint2(0u, 0u) : make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
          }
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(a1_full, 0);
    if (ew == 0) LQ_STAMP(3);

    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint8_t* sb = staging + ew * 2048;
    const uint32_t row_base = (uint32_t)lane * 64u, swz = ((uint32_t)lane >> 1) & 3u;   // 64-byte rows, SWIZZLE_64B
    const int m0 = row0 + quarter * 32;
#pragma unroll 1
    for (int nc = 0; nc < n_chunks; ++nc) {
      const int b = nc & 1, u = nc >> 1;
      mbar_wait(&acc_full[b], u & 1);
      tc_fence_after();
      if (ew == 0) LQ_STAMP(48 + 2 * nc);
#pragma unroll 1
      for (int c = part; c < 8; c += 4) {
        const int n0 = nc * 256 + c * 32;
        if (n0 >= p.N) break;                                   // warp-uniform; N % 32 == 0
        uint32_t rg[32];
        tmem_ld_32x32(tl + b * 256 + c * 32, rg);
        float4 bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          bv[i] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        tmem_ld_wait();
        if (lane == 0) bulk_wait_read<0>();                     // this warp's previous store has left the staging buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(sb + row_base + (((uint32_t)j ^ swz) << 4)) =
              make_uint4(pack_bf16x2(__uint_as_float(rg[8 * j]) + bv[2 * j].x, __uint_as_float(rg[8 * j + 1]) + bv[2 * j].y),
                         pack_bf16x2(__uint_as_float(rg[8 * j + 2]) + bv[2 * j].z, __uint_as_float(rg[8 * j + 3]) + bv[2 * j].w),
                         pack_bf16x2(__uint_as_float(rg[8 * j + 4]) + bv[2 * j + 1].x, __uint_as_float(rg[8 * j + 5]) + bv[2 * j + 1].y),
                         pack_bf16x2(__uint_as_float(rg[8 * j + 6]) + bv[2 * j + 1].z, __uint_as_float(rg[8 * j + 7]) + bv[2 * j + 1].w));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_out, sb, n0, m0);                  // rows past M are clipped by the TMA unit
          bulk_commit();
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&acc_empty[b], 0);
      if (ew == 0) LQ_STAMP(48 + 2 * nc + 1);
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
    if (ew == 0) LQ_STAMP(4);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

template <int C>
static int launch_ln_qkv(const LnQkvParams& p, const void* w, void* out, cudaStream_t stream) {
  using P = LnQkvPlan<C>;
  static unsigned long long attr_devices = 0;
  if (first_use_on_device(attr_devices)) {
    cudaError_t e = cudaFuncSetAttribute(swin_ln_qkv_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM);
    if (e != cudaSuccess) return (int)e;
  }
  CUtensorMap tw, to;
  int rc;
  if ((rc = make_tmap(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, p.N, C, C, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, p.M, p.N, p.N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE)) != MVLT_OK) return rc;
  const int pairs = (p.M + 255) / 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(LQ_THREADS);
  cfg.dynamicSmemBytes = P::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mvlt_pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, swin_ln_qkv_kernel<C>, tw, to, p);
  return e == cudaSuccess ? MVLT_OK : (int)e;
}

}  // namespace mvlt

using namespace mvlt;

// debug hook (not part of include/mvlt_b200.h)
extern "C" int mvlt_debug_lnqkv_trace(void* dev_buf) {
  g_lnqkv_trace = reinterpret_cast<unsigned long long*>(dev_buf);
  return MVLT_OK;
}

// qkv (bf16 [B*H*W, N], rows WINDOW-MAJOR for the image rolled by -shift, dense) = LayerNorm(x) . w^T + bias in one kernel.
// x: fp32 [B*H*W, C] natural token order (row stride ldx); w: bf16 [N, C] (nn.Linear layout), bias fp32 [N] or NULL; N % 32 == 0.
// C in {96, 192, 384}.  Replaces vfe.py:356 + :361-364 + :231 (norm1, roll, window_partition, qkv) of one SwinTransformerBlock; the
// output is what mvlt_window_attention_tc reads.
extern "C" int mvlt_swin_ln_qkv(const float* x, long long ldx, const float* gamma, const float* beta, float eps, const void* w,
                                const float* bias, void* out, int B, int H, int W, int C, int N, int window, int shift,
                                cudaStream_t stream) {
  if (!x || !gamma || !beta || !w || !out || B <= 0 || H <= 0 || W <= 0 || N <= 0) return MVLT_ERR_INVALID;
  if ((C != 96 && C != 192 && C != 384) || N % 32 != 0) return MVLT_ERR_UNSUPPORTED;
  if (window != LQ_WS) return MVLT_ERR_UNSUPPORTED;
  if (H % window || W % window || shift < 0 || shift >= window) return MVLT_ERR_INVALID;
  if (ldx < C || ldx % 4 != 0) return MVLT_ERR_INVALID;
  if (((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)out & 15) || ((uintptr_t)gamma & 15) || ((uintptr_t)beta & 15) ||
      (bias && ((uintptr_t)bias & 15))) return MVLT_ERR_INVALID;
  const long long M = (long long)B * H * W;
  if (M > 0x7fffffffLL - 512) return MVLT_ERR_UNSUPPORTED;
  int rc = mvlt_gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  LnQkvParams p;
  p.x = x; p.ldx = ldx; p.gamma = gamma; p.beta = beta; p.bias = bias; p.eps = eps; p.M = (int)M; p.N = N;
  p.H = H; p.W = W; p.shift = shift; p.nWw = W / window; p.nW = (H / window) * (W / window);
  auto log2_exact = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  p.nW_shift = log2_exact(p.nW); p.nWw_shift = log2_exact(p.nWw);
  if (p.nW_shift < 0 || p.nWw_shift < 0) p.nW_shift = p.nWw_shift = -1;
  p.trace = g_lnqkv_trace;
  return C == 384 ? launch_ln_qkv<384>(p, w, out, stream) : (C == 192 ? launch_ln_qkv<192>(p, w, out, stream) : launch_ln_qkv<96>(p, w, out, stream));
}
