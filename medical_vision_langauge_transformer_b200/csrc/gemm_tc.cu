// tcgen05 / TMEM / TMA GEMM for every nn.Linear on the MVLT hot path (bf16 operands, fp32 accumulate):
//
//     C[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual        (nn.Linear weight layout == K-major B operand)
//
// replaces: vfe.py:231 (qkv), :252 (proj), :136-139 (fc1/fc2), :443 (merge reduction);
//           HF modeling_bert.py:179-181 (Q,K,V as one [2304,768] weight), :295, :338, :352, :476, :502 (MLM decoder).
//
// One persistent, warp-specialised kernel, templated on the CTA-group size CG:
//   CG = 2 (default): a cluster of two CTAs (one SM pair) owns a 256 x BLOCK_N output tile.  Each CTA TMA-loads its
//           own 128 rows of A and HALF of the W tile per 64-wide k-block; one thread of the leader CTA issues
//           tcgen05.mma.cta_group::2 (M=256) which reads both CTAs' smem and writes both CTAs' TMEM.  Per SM this
//           needs 32 KB of operands per 512 tensor cycles instead of 48 KB per 512 -> 6-stage ring, ~3x less
//           L2->SM traffic per FLOP than the single-CTA tile.
//   CG = 1: single-CTA 128 x BLOCK_N tiles (kept for A/B measurements: MVLT_GEMM_CTAS=1).
// Roles per CTA (384 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator,
// warps 4-11 epilogue.  Two TMEM accumulator stages (2 x 256 columns): the epilogue of tile i overlaps the MMAs of
// tile i+1.  Epilogue: tcgen05.ld 32x32b.x32 -> +bias (smem-staged) -> GELU/tanh -> XOR-swizzled smem transpose ->
// +residual -> 128 B-per-row coalesced stores.  BLOCK_N is a runtime value (multiple of 32, <= 256) carried by the
// instruction descriptor and the TMA box.  M/N/K edges: TMA zero fill on loads, guards on stores.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"

namespace mvlt {

constexpr int BM = 128;  // rows per CTA
constexpr int BK = 64;   // 64 bf16 = 128 B = one swizzle span
constexpr int BN_MAX = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int RING_BYTES = 192 * 1024;
constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * 32 * 32 * 4;  // one swizzled 32x32 fp32 tile per epilogue warp
constexpr int AUX_BYTES = 256 /*barriers*/ + BN_MAX * 4 /*bias*/;
constexpr int GEMM_SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + AUX_BYTES + EPI_STAGING_BYTES;

template <int CG> struct Cfg {
  static constexpr int B_STAGE_BYTES = (BN_MAX / CG) * BK * 2;  // W rows held per CTA per stage
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = RING_BYTES / STAGE_BYTES;  // 4 (CG=1) or 6 (CG=2)
};

struct GemmParams {
  void* C;
  long long ldc;
  const float* bias;
  const void* res;
  long long ldres;
  int M, N, K;
  int block_n;
  int act;        // 0 none, 1 erf-GELU, 2 tanh
  int out_dtype;  // MVLT_F32 / MVLT_BF16
  int res_dtype;  // -1 none, MVLT_F32, MVLT_BF16
  int tiles_m, tiles_n;
  int debug;      // MVLT_GEMM_DEBUG bits (profiling experiments only): 1 no epilogue global traffic, 2 no MMA, 4 no TMA
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return gelu_erf_fast(v);
  if (act == 2) return tanhf(v);
  return v;
}

// ---- cta_group-templated PTX -----------------------------------------------------------------------------------
template <int CG> __device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
  else asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
}
template <int CG> __device__ __forceinline__ void tmem_relinquish_cg() {
  if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG> __device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_bf16_cg(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive once all MMAs issued so far by this thread retire; CG=2: on the same barrier of BOTH CTAs of the pair
template <int CG> __device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMA tile load into THIS CTA's smem; CG=2: completion bytes are counted on the LEADER CTA's mbarrier (peer bit cleared)
template <int CG>
__device__ __forceinline__ void tma_load_cg(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  if (CG == 1) {
    tma_load_2d(smem_dst, tmap, bar, c0, c1);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmParams p) {
  using C = Cfg<CG>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]  TMA -> MMA        (lives in the leader CTA)
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]  MMA -> TMA        (one copy per CTA)
  uint64_t* tmem_full = bars + 2 * STAGES;       // [2]       MMA -> epilogue   (one copy per CTA)
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]       epilogue -> MMA   (leader CTA)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* bias_s = reinterpret_cast<float*>(smem + RING_BYTES + 256);
  float* staging = reinterpret_cast<float*>(smem + RING_BYTES + AUX_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int group = blockIdx.x / CG, num_groups = gridDim.x / CG;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int b_rows = p.block_n / CG;  // W rows this CTA loads per stage

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
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
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_cg<CG>(tmem_ptr, TMEM_COLS);
    tmem_relinquish_cg<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

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
      for (int kb = 0; kb < num_kb; ++kb, ++kc) {
        const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          if (p.debug & 4) {
            if (rank == 0) mbar_arrive(&full_bar[s]);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
            tma_load_cg<CG>(smem_a + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BK, m0);
            tma_load_cg<CG>(smem_b + s * C::B_STAGE_BYTES, &tmap_b, &full_bar[s], kb * BK, n0);
          }
        }
        __syncwarp();
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
        const uint32_t tmem_d = tmem_base + acc * BN_MAX;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t lo_a = lo_a0 + s * (A_STAGE_BYTES >> 4);
          const uint32_t lo_b = lo_b0 + s * (C::B_STAGE_BYTES >> 4);
          const int ksteps = min(BK, p.K - kb * BK) / 16;  // K % 16 == 0 is checked on the host
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements = 32 B inside the 128 B swizzle span: +2 in the (addr >> 4) field
              if (k < ksteps && !(p.debug & 2))
                umma_bf16_cg<CG>(tmem_d, ((uint64_t)DESC_HI << 32) | (lo_a + 2 * k), ((uint64_t)DESC_HI << 32) | (lo_b + 2 * k),
                                 idesc, (kb | k) != 0);
            }
            umma_commit_cg<CG>(&empty_bar[s]);  // smem slot reusable (in both CTAs) once these MMAs retire
            if (kb == num_kb - 1) umma_commit_cg<CG>(&tmem_full[acc]);  // accumulator complete (both CTAs' epilogues)
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------- epilogue (every CTA, its own 128 rows) -----------------
    const int ew = warp - EPI_WARP0;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the only ones this warp may touch
    const int half = ew >> 2;      // column half of the tile
    const int chunks = p.block_n / 32;
    const int c_begin = half * ((chunks + 1) / 2);
    const int c_end = half ? chunks : (chunks + 1) / 2;
    const bool vec_ok = (p.ldc % 4 == 0) && (p.res_dtype < 0 || p.ldres % 4 == 0);
    float* stg = staging + ew * 1024;
    const int crow = lane >> 3, cch = lane & 7;  // coalesced layout: row-in-group and 16 B column chunk of this lane
    uint32_t it = 0;
    for (int tile = group; tile < num_tiles; tile += num_groups, ++it) {
      const uint32_t acc = it & 1, acc_ph = (it >> 1) & 1;
      const int m0 = (tile / p.tiles_n) * (BM * CG) + rank * BM + quarter * 32;
      const int n0 = (tile % p.tiles_n) * p.block_n;
      // stage this tile's bias slice in smem while the MMAs run
      epi_bar_sync();  // everyone is done reading the previous tile's slice
      if (et < p.block_n) bias_s[et] = (p.bias && n0 + et < p.N) ? __ldg(p.bias + n0 + et) : 0.f;
      epi_bar_sync();
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      for (int c = c_begin; c < c_end; ++c) {
        const int nb = n0 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN_MAX + c * 32, r);
        tmem_ld_wait();
        if (nb >= p.N) continue;  // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
          float4 v = make_float4(__uint_as_float(r[j]) + b.x, __uint_as_float(r[j + 1]) + b.y,
                                 __uint_as_float(r[j + 2]) + b.z, __uint_as_float(r[j + 3]) + b.w);
          if (p.act) {
            v.x = apply_act(v.x, p.act); v.y = apply_act(v.y, p.act);
            v.z = apply_act(v.z, p.act); v.w = apply_act(v.w, p.act);
          }
          *reinterpret_cast<float4*>(stg + lane * 32 + ((((j >> 2) ^ (lane & 7))) << 2)) = v;
        }
        __syncwarp();
        const int n = nb + cch * 4;
        const bool vec = vec_ok && n + 4 <= p.N;
        float4 v[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int row = g * 4 + crow;
          v[g] = *reinterpret_cast<const float4*>(stg + row * 32 + ((cch ^ (row & 7)) << 2));
        }
        if (!(p.debug & 1) && n < p.N) {
          if (vec) {
            // all residual loads first (the in-place residual aliases C, so the compiler may not hoist them itself)
            if (p.res_dtype == MVLT_F32) {
              float4 t[8];
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const int m = m0 + g * 4 + crow;
                t[g] = m < p.M ? load4(reinterpret_cast<const float*>(p.res) + (long long)m * p.ldres + n) : make_float4(0, 0, 0, 0);
              }
#pragma unroll
              for (int g = 0; g < 8; ++g) { v[g].x += t[g].x; v[g].y += t[g].y; v[g].z += t[g].z; v[g].w += t[g].w; }
            } else if (p.res_dtype == MVLT_BF16) {
              float4 t[8];
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const int m = m0 + g * 4 + crow;
                t[g] = m < p.M ? load4(reinterpret_cast<const bf16*>(p.res) + (long long)m * p.ldres + n) : make_float4(0, 0, 0, 0);
              }
#pragma unroll
              for (int g = 0; g < 8; ++g) { v[g].x += t[g].x; v[g].y += t[g].y; v[g].z += t[g].z; v[g].w += t[g].w; }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int m = m0 + g * 4 + crow;
              if (m >= p.M) continue;
              if (p.out_dtype == MVLT_F32) store4(reinterpret_cast<float*>(p.C) + (long long)m * p.ldc + n, v[g]);
              else store4(reinterpret_cast<bf16*>(p.C) + (long long)m * p.ldc + n, v[g]);
            }
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int m = m0 + g * 4 + crow;
              if (m >= p.M) continue;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (n + q >= p.N) break;
                float x = q == 0 ? v[g].x : q == 1 ? v[g].y : q == 2 ? v[g].z : v[g].w;
                const long long ro = (long long)m * p.ldres + n + q, co = (long long)m * p.ldc + n + q;
                if (p.res_dtype == MVLT_F32) x += reinterpret_cast<const float*>(p.res)[ro];
                else if (p.res_dtype == MVLT_BF16) x += to_f32(reinterpret_cast<const bf16*>(p.res)[ro]);
                if (p.out_dtype == MVLT_F32) reinterpret_cast<float*>(p.C)[co] = x;
                else reinterpret_cast<bf16*>(p.C)[co] = __float2bfloat16_rn(x);
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_remote(&tmem_empty[acc], 0);
        else mbar_arrive(&tmem_empty[acc]);
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int g_num_sms = 0;
static int g_ctas = 2;
static int g_debug = 0;

static int gemm_tc_init() {
  if (g_encode) return MVLT_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return MVLT_ERR_DRIVER;
  e = cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (const char* s = getenv("MVLT_GEMM_CTAS")) g_ctas = atoi(s) == 1 ? 1 : 2;
  if (const char* s = getenv("MVLT_GEMM_DEBUG")) g_debug = atoi(s);
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return MVLT_OK;
}

static int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld_elems, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVLT_OK : MVLT_ERR_DRIVER;
}

// Tile-width heuristic: maximise (useful columns / padded columns) x (wave quantisation) x (smem-feed efficiency).
static int pick_block_n(int M, int N, int groups, int cg) {
  const int cand[] = {256, 192, 128, 96, 64, 32};
  const int tiles_m = (M + BM * cg - 1) / (BM * cg);
  double best = -1;
  int best_bn = 128;
  for (int bn : cand) {
    const int tn = (N + bn - 1) / bn;
    const double useful = (double)N / ((double)tn * bn);
    const long long tiles = (long long)tiles_m * tn;
    const long long waves = (tiles + groups - 1) / groups;
    const double wave_eff = (double)tiles / ((double)waves * groups);
    // tensor cycles per K=16 step vs cycles to read this CTA's operand slices from smem at 128 B/clk
    const double mma_cycles = bn / 2.0;
    const double feed = mma_cycles / ((4096.0 + (bn / cg) * 32.0) / 128.0);
    const double score = useful * wave_eff * (feed < 1.0 ? feed : 1.0);
    if (score > best + 1e-9) {
      best = score;
      best_bn = bn;
    }
  }
  return best_bn;
}

template <int CG>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<CG>, ta, tb, p);
  return e == cudaSuccess ? MVLT_OK : (int)e;
}

}  // namespace mvlt

using namespace mvlt;

extern "C" int mvlt_gemm_tc_init(void) { return gemm_tc_init(); }

extern "C" int mvlt_gemm_bf16_tc(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc,
                                 const float* bias, const void* residual, long long ldres, int res_dtype, int M, int N,
                                 int K, int act, int out_dtype, int block_n, cudaStream_t stream) {
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0) return MVLT_ERR_INVALID;
  if (K % 16 != 0 || lda % 8 != 0 || ldw % 8 != 0) return MVLT_ERR_INVALID;  // TMA: 16 B aligned rows
  if (((uintptr_t)A & 15) || ((uintptr_t)W & 15)) return MVLT_ERR_INVALID;
  if (act < 0 || act > 2 || (out_dtype != MVLT_F32 && out_dtype != MVLT_BF16)) return MVLT_ERR_INVALID;
  if (!residual) res_dtype = -1;
  if (res_dtype > MVLT_BF16) return MVLT_ERR_INVALID;
  int rc = gemm_tc_init();
  if (rc != MVLT_OK) return rc;
  const int cg = g_ctas;
  const int groups = g_num_sms / cg;
  if (block_n <= 0) block_n = pick_block_n(M, N, groups, cg);
  if (block_n % 32 != 0 || block_n > BN_MAX) return MVLT_ERR_INVALID;

  CUtensorMap ta, tb;
  if ((rc = make_tmap(&ta, A, M, K, lda, BM)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tb, W, N, K, ldw, block_n / cg)) != MVLT_OK) return rc;

  GemmParams p;
  p.C = C; p.ldc = ldc; p.bias = bias; p.res = residual; p.ldres = ldres;
  p.M = M; p.N = N; p.K = K; p.block_n = block_n; p.act = act; p.out_dtype = out_dtype; p.res_dtype = res_dtype;
  p.tiles_m = (M + BM * cg - 1) / (BM * cg);
  p.tiles_n = (N + block_n - 1) / block_n;
  p.debug = g_debug;
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = cg * (tiles < groups ? tiles : groups);
  return cg == 2 ? launch<2>(ta, tb, p, grid, stream) : launch<1>(ta, tb, p, grid, stream);
}
