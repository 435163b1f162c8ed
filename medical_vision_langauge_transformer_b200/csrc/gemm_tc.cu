// tcgen05 / TMEM / TMA GEMM for every nn.Linear on the MVLT hot path (bf16 operands, fp32 accumulate):
//
//     C[M,N] = epilogue( A[M,K] . W[N,K]^T )          (nn.Linear weight layout == K-major B operand)
//
// replaces: vfe.py:231 (qkv), :252 (proj), :136-139 (fc1/fc2), :443 (merge reduction);
//           HF modeling_bert.py:179-181 (Q,K,V as one [2304,768] weight), :295, :338, :352, :463 (pooler).
//
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0    : TMA producer   — cp.async.bulk.tensor 2-D boxes {64 x 128} of A and {64 x BLOCK_N} of W into a
//                                4-stage smem ring (128B swizzle), completion on mbarriers
//   warp 1    : MMA issuer     — one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16),
//                                accumulating into one of two TMEM accumulator stages (2 x 256 columns)
//   warp 2    : TMEM allocator
//   warps 4-11: epilogue       — tcgen05.ld 32x32b.x32 -> registers -> +bias -> GELU/tanh -> smem transpose ->
//                                +residual -> coalesced 128 B row stores; runs on accumulator stage i while the MMA
//                                warp fills stage i^1
// BLOCK_N is a runtime value (multiple of 32, <= 256) carried in the instruction descriptor and the TMA box, so one
// kernel serves N = 96 ... 3072.  M/N/K edges are handled by TMA zero fill on loads and guards on stores.
#include <cuda.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace mvlt {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int BN_MAX = 256;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = BN_MAX * BK * 2;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * 32 * 32 * 4;  // one swizzled 32x32 fp32 tile per epilogue warp
constexpr int GEMM_SMEM_BYTES =
    STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGING_BYTES;

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
  int debug;  // MVLT_GEMM_DEBUG bits (profiling experiments only): 1 = no epilogue global traffic, 2 = no MMA, 4 = no TMA
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return gelu_erf(v);
  if (act == 2) return tanhf(v);
  return v;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]       MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

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
      mbar_init(&tmem_empty[s], NUM_EPI_WARPS * 32);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)p.block_n * BK * 2;
      uint32_t kc = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * p.block_n;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (p.debug & 4) { mbar_arrive(&full_bar[s]); continue; }
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BK, m0);
          tma_load_2d(smem_b + s * B_STAGE_BYTES, &tmap_b, &full_bar[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, (uint32_t)p.block_n);
      uint32_t kc = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN_MAX;
        for (int kb = 0; kb < num_kb; ++kb, ++kc) {
          const uint32_t s = kc % STAGES, ph = (kc / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(base + s * A_STAGE_BYTES);
          const uint64_t db = umma_desc_k_sw128(base + STAGES * A_STAGE_BYTES + s * B_STAGE_BYTES);
          const int ksteps = min(BK, p.K - kb * BK) / 16;  // K % 16 == 0 is checked on the host
#pragma unroll 1
          for (int k = 0; k < ksteps && !(p.debug & 2); ++k) {
            // advance 16 elements = 32 B inside the 128 B swizzle span: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs retire
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------- epilogue -----------------------------------
    // TMEM gives each thread one ROW of the 32x32 chunk; writing rows straight to global would touch 32 different
    // 128 B lines per instruction.  So: bias + activation in the row layout, transpose through a per-warp 4 KB
    // XOR-swizzled smem tile (conflict-free both ways), then residual add + convert + store in the COALESCED layout
    // (8 lanes cover one row's 128 B, 4 rows per instruction).
    const int ew = warp - EPI_WARP0;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the only ones this warp may touch
    const int half = ew >> 2;      // column half of the tile
    const int chunks = p.block_n / 32;
    const int c_begin = half * ((chunks + 1) / 2);
    const int c_end = half ? chunks : (chunks + 1) / 2;
    const bool vec_ok = (p.ldc % 4 == 0) && (p.res_dtype < 0 || p.ldres % 4 == 0);
    float* stg = reinterpret_cast<float*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256) + ew * 1024;
    const int crow = lane >> 3, cch = lane & 7;  // coalesced layout: this lane's row-in-group and 16 B column chunk
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_ph = (it >> 1) & 1;
      const int m0 = (tile / p.tiles_n) * BM + quarter * 32;
      const int n0 = (tile % p.tiles_n) * p.block_n;
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      for (int c = c_begin; c < c_end; ++c) {
        const int nb = n0 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN_MAX + c * 32, r);
        tmem_ld_wait();
        if (nb >= p.N) continue;  // warp-uniform
        const bool full_n = nb + 32 <= p.N;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                 __uint_as_float(r[j + 3]));
          if (p.bias) {
            if (full_n) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            } else {
              if (nb + j < p.N) v.x += p.bias[nb + j];
              if (nb + j + 1 < p.N) v.y += p.bias[nb + j + 1];
              if (nb + j + 2 < p.N) v.z += p.bias[nb + j + 2];
              if (nb + j + 3 < p.N) v.w += p.bias[nb + j + 3];
            }
          }
          if (p.act) {
            v.x = apply_act(v.x, p.act); v.y = apply_act(v.y, p.act);
            v.z = apply_act(v.z, p.act); v.w = apply_act(v.w, p.act);
          }
          *reinterpret_cast<float4*>(stg + lane * 32 + ((((j >> 2) ^ (lane & 7))) << 2)) = v;
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int row = g * 4 + crow;
          const int m = m0 + row, n = nb + cch * 4;
          float4 v = *reinterpret_cast<const float4*>(stg + row * 32 + ((cch ^ (row & 7)) << 2));
          if (m >= p.M || n >= p.N || (p.debug & 1)) continue;
          if (vec_ok && n + 4 <= p.N) {
            if (p.res_dtype == MVLT_F32) {
              const float4 t = load4(reinterpret_cast<const float*>(p.res) + (long long)m * p.ldres + n);
              v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            } else if (p.res_dtype == MVLT_BF16) {
              const float4 t = load4(reinterpret_cast<const bf16*>(p.res) + (long long)m * p.ldres + n);
              v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            if (p.out_dtype == MVLT_F32) store4(reinterpret_cast<float*>(p.C) + (long long)m * p.ldc + n, v);
            else store4(reinterpret_cast<bf16*>(p.C) + (long long)m * p.ldc + n, v);
          } else {
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (n + q >= p.N) break;
              float x = e[q];
              const long long ro = (long long)m * p.ldres + n + q, co = (long long)m * p.ldc + n + q;
              if (p.res_dtype == MVLT_F32) x += reinterpret_cast<const float*>(p.res)[ro];
              else if (p.res_dtype == MVLT_BF16) x += to_f32(reinterpret_cast<const bf16*>(p.res)[ro]);
              if (p.out_dtype == MVLT_F32) reinterpret_cast<float*>(p.C)[co] = x;
              else reinterpret_cast<bf16*>(p.C)[co] = __float2bfloat16_rn(x);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int g_num_sms = 0;

static int gemm_tc_init() {
  if (g_encode) return MVLT_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return MVLT_ERR_DRIVER;
  e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
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
static int pick_block_n(int M, int N, int sms) {
  const int cand[] = {256, 192, 128, 96, 64, 32};
  const int tiles_m = (M + BM - 1) / BM;
  double best = -1;
  int best_bn = 128;
  for (int bn : cand) {
    const int tn = (N + bn - 1) / bn;
    const double useful = (double)N / ((double)tn * bn);
    const long long tiles = (long long)tiles_m * tn;
    const long long waves = (tiles + sms - 1) / sms;
    const double wave_eff = (double)tiles / ((double)waves * sms);
    const double feed = (bn / 2.0) / ((4096.0 + bn * 32.0) / 128.0);  // MMA cycles / smem-read cycles
    const double score = useful * wave_eff * (feed < 1.0 ? feed : 1.0);
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
  if (block_n <= 0) block_n = pick_block_n(M, N, g_num_sms);
  if (block_n % 32 != 0 || block_n > BN_MAX) return MVLT_ERR_INVALID;

  CUtensorMap ta, tb;
  if ((rc = make_tmap(&ta, A, M, K, lda, BM)) != MVLT_OK) return rc;
  if ((rc = make_tmap(&tb, W, N, K, ldw, block_n)) != MVLT_OK) return rc;

  GemmParams p;
  p.C = C; p.ldc = ldc; p.bias = bias; p.res = residual; p.ldres = ldres;
  p.M = M; p.N = N; p.K = K; p.block_n = block_n; p.act = act; p.out_dtype = out_dtype; p.res_dtype = res_dtype;
  p.tiles_m = (M + BM - 1) / BM;
  p.tiles_n = (N + block_n - 1) / block_n;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("MVLT_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_tc_kernel<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(ta, tb, p);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
