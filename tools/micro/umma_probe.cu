// Hardware probes for the tcgen05 features the attention kernels (csrc/attention_tc.cu) rely on, each checked against a CPU
// product.  One CTA per probe, operands written to shared memory by plain stores with the swizzle applied by hand.
//   P1  SS MMA, both operands K-major SWIZZLE_64B (64-byte rows: head_dim 32), M=128 N=128 K=32
//   P2  SS MMA, B operand MN-major SWIZZLE_64B, N=64 taken from two 32-wide atoms LBO bytes apart, K=64 (A K-major SW128)
//   P3  SS MMA, B operand MN-major SWIZZLE_128B (128-byte rows: head_dim 64), N=64, K=144, M=128
//   P4  TS MMA: A operand (bf16 pairs) written to TMEM with tcgen05.st, same B as P2
//   P5  disable-output-lane mask: second accumulate pass restricted to lanes 0..63
//   P6  tcgen05.ld 32x32b .x16 / .x1 at an odd column, N=144 accumulator
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I ../../medical_vision_langauge_transformer_b200/csrc umma_probe.cu -o umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "common.cuh"

using namespace mvlt;

int mvlt_pdl_enabled(void) { return 0; }

static __host__ __device__ inline float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// descriptor helpers under test -------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_make(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
constexpr uint32_t L_SW128 = 2, L_SW64 = 4;
__device__ __forceinline__ uint32_t idesc_make(uint32_t m, uint32_t n, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_mask(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc, uint32_t m0, uint32_t m1,
                                            uint32_t m2, uint32_t m3) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

// shared-memory images ----------------------------------------------------------------------------------------------
// K-major / row-major tile with 64-byte rows (32 bf16), SWIZZLE_64B: 16 B chunk j of row r sits at chunk j ^ ((r >> 1) & 3)
__device__ __forceinline__ uint32_t off_sw64(int r, int c) { return r * 64 + ((((c >> 3) ^ ((r >> 1) & 3))) << 4) + (c & 7) * 2; }
// 128-byte rows (64 bf16), SWIZZLE_128B: chunk j -> j ^ (r & 7)
__device__ __forceinline__ uint32_t off_sw128(int r, int c) { return r * 128 + ((((c >> 3) ^ (r & 7))) << 4) + (c & 7) * 2; }

struct ProbeArgs {
  const bf16* a;   // A source, row-major
  const bf16* b;   // B source, row-major
  float* out;      // [128][ncols]
  int which;
  int variant;   // 1: LBO / SBO of the MN-major B descriptor swapped
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint8_t* sa = smem;            // A image
  uint8_t* sb = smem + (p.which == 3 ? 65536 : 32768);    // B image
  for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  int ncols = 0;
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  if (p.which == 1) {
    // A = Q [128][32], B = K [128][32], both SW64 K-major
    for (int i = tid; i < 128 * 32; i += 128) {
      const int r = i / 32, c = i % 32;
      *reinterpret_cast<bf16*>(sa + off_sw64(r, c)) = p.a[i];
      *reinterpret_cast<bf16*>(sb + off_sw64(r, c)) = p.b[i];
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      const uint32_t id = idesc_make(128, 128, 0);
      for (int k = 0; k < 2; ++k)
        umma_bf16(tmem, desc_make(smem_u32(sa) + 32 * k, 16, 512, L_SW64), desc_make(smem_u32(sb) + 32 * k, 16, 512, L_SW64), id, k);
      umma_commit(&bar);
    }
    ncols = 128;
  } else if (p.which == 2 || p.which == 4 || p.which == 5) {
    // A = P [128][64 keys] K-major SW128; B = V_A [64 keys][32 d] at sb, V_B at sb + 4096, SW64 rows, MN-major, N = 64
    for (int i = tid; i < 128 * 64; i += 128) {
      const int r = i / 64, c = i % 64;
      *reinterpret_cast<bf16*>(sa + off_sw128(r, c)) = p.a[i];
    }
    for (int i = tid; i < 64 * 64; i += 128) {   // p.b is [64 keys][64 n]: n < 32 -> V_A, n >= 32 -> V_B
      const int key = i / 64, n = i % 64;
      *reinterpret_cast<bf16*>(sb + (n >> 5) * 4096 + off_sw64(key, n & 31)) = p.b[i];
    }
    fence_proxy_async_smem();
    __syncthreads();
    const uint32_t id = idesc_make(128, 64, 1);
    if (p.which == 4) {
      // P as bf16 pairs into TMEM columns [256, 288): thread = row, register j = keys (2j, 2j+1)
      uint32_t r[32];
      const int row = warp * 32 + lane;
      for (int j = 0; j < 32; ++j) {
        const __nv_bfloat162 t = __halves2bfloat162(p.a[row * 64 + 2 * j], p.a[row * 64 + 2 * j + 1]);
        r[j] = *reinterpret_cast<const uint32_t*>(&t);
      }
      tmem_st_32x32(tl + 256, r);
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        for (int k = 0; k < 4; ++k) mma_ts(tmem, tmem + 256 + 8 * k, desc_make(smem_u32(sb) + 1024 * k, p.variant ? 512 : 4096, p.variant ? 4096 : 512, L_SW64), id, k);
        umma_commit(&bar);
      }
    } else if (tid == 0) {
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, desc_make(smem_u32(sa) + 32 * k, 16, 1024, L_SW128), desc_make(smem_u32(sb) + 1024 * k, p.variant ? 512 : 4096, p.variant ? 4096 : 512, L_SW64), id, k);
      if (p.which == 5)   // second pass, lanes 64..127 disabled: rows 0..63 end up doubled
        for (int k = 0; k < 4; ++k)
          mma_ss_mask(tmem, desc_make(smem_u32(sa) + 32 * k, 16, 1024, L_SW128), desc_make(smem_u32(sb) + 1024 * k, p.variant ? 512 : 4096, p.variant ? 4096 : 512, L_SW64), id, 1,
                      0u, 0u, 0xffffffffu, 0xffffffffu);
      umma_commit(&bar);
    }
    ncols = 64;
  } else if (p.which == 3 || p.which == 6) {
    // A = P [128][144 keys] as three K-major SW128 k-blocks of 64 keys (16 KB each; the last one half used);
    // B = V [144 keys][64 d] SW128 rows, MN-major, N = 64  (which 6: B = K [144][64] K-major, N = 144, K = 64)
    if (p.which == 3) {
      for (int i = tid; i < 128 * 144; i += 128) {
        const int r = i / 144, c = i % 144;
        *reinterpret_cast<bf16*>(sa + (c >> 6) * 16384 + off_sw128(r, c & 63)) = p.a[i];
      }
      for (int i = tid; i < 144 * 64; i += 128) *reinterpret_cast<bf16*>(sb + off_sw128(i / 64, i % 64)) = p.b[i];
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        const uint32_t id = idesc_make(128, 64, 1);
        for (int k = 0; k < 9; ++k)
          umma_bf16(tmem, desc_make(smem_u32(sa) + (k >> 2) * 16384 + 32 * (k & 3), 16, 1024, L_SW128),
                    desc_make(smem_u32(sb) + 2048 * k, p.variant ? 1024 : 16, p.variant ? 16 : 1024, L_SW128), id, k);
        umma_commit(&bar);
      }
      ncols = 64;
    } else {
      for (int i = tid; i < 128 * 64; i += 128) *reinterpret_cast<bf16*>(sa + off_sw128(i / 64, i % 64)) = p.a[i];
      for (int i = tid; i < 144 * 64; i += 128) *reinterpret_cast<bf16*>(sb + off_sw128(i / 64, i % 64)) = p.b[i];
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        const uint32_t id = idesc_make(128, 144, 0);
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, desc_make(smem_u32(sa) + 32 * k, 16, 1024, L_SW128), desc_make(smem_u32(sb) + 32 * k, 16, 1024, L_SW128), id, k);
        umma_commit(&bar);
      }
      ncols = 144;
    }
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  if (p.which == 6) {
    // 144 columns read as x32 x4, then x16 at column 128; plus x1 reads at odd columns 49 and 131 appended as cols 144, 145
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tl + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.out[row * 146 + c0 + j] = __uint_as_float(r[j]);
    }
    uint32_t r16[16];
    tmem_ld_x16(tl + 128, r16);
    uint32_t r1a, r1b;
    tmem_ld_x1(tl + 49, r1a);
    tmem_ld_x1(tl + 131, r1b);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) p.out[row * 146 + 128 + j] = __uint_as_float(r16[j]);
    p.out[row * 146 + 144] = __uint_as_float(r1a);
    p.out[row * 146 + 145] = __uint_as_float(r1b);
  } else {
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tl + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.out[row * ncols + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static std::vector<bf16> rnd(int n, unsigned seed) {
  std::vector<bf16> v(n);
  unsigned s = seed * 2654435761u + 12345u;
  for (int i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    v[i] = __float2bfloat16_rn(((int)((s >> 9) & 0xffff) - 32768) / 32768.0f);
  }
  return v;
}
static float f(const bf16& x) { return __bfloat162float(x); }

static int run(int which, const std::vector<bf16>& a, const std::vector<bf16>& b, const std::vector<float>& ref, int ncols, const char* name, int variant = 0) {
  bf16 *da, *db;
  float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * 160 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 128 * 160 * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  ProbeArgs p{da, db, dout, which, variant};
  probe_kernel<<<1, 128, 100 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("P%d %-44s LAUNCH ERROR %s\n", which, name, cudaGetErrorString(e)); return 1; }
  std::vector<float> out(128 * ncols);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  int bad = 0, first = -1;
  for (int i = 0; i < 128 * ncols; ++i) {
    const double d = fabs(out[i] - ref[i]);
    maxerr = fmax(maxerr, d); maxref = fmax(maxref, fabs(ref[i]));
    if (d > 1e-2 * fmax(1.0, fabs(ref[i]))) { ++bad; if (first < 0) first = i; }
  }
  printf("P%d %-44s %s  max|err| %.3e  max|ref| %.3f  bad %d/%d", which, name, bad ? "FAIL" : "PASS", maxerr, maxref, bad, 128 * ncols);
  if (bad) printf("  first bad (row %d, col %d): got %.4f want %.4f", first / ncols, first % ncols, out[first], ref[first]);
  printf("\n");
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return bad != 0;
}

int main() {
  int fails = 0;
  {  // P1
    auto q = rnd(128 * 32, 1), k = rnd(128 * 32, 2);
    std::vector<float> ref(128 * 128);
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) { float s = 0; for (int d = 0; d < 32; ++d) s += f(q[i * 32 + d]) * f(k[j * 32 + d]); ref[i * 128 + j] = s; }
    fails += run(1, q, k, ref, 128, "SS K-major SW64 x SW64, N=128 K=32");
  }
  auto p64 = rnd(128 * 64, 3), v64 = rnd(64 * 64, 4);
  std::vector<float> ref2(128 * 64);
  for (int i = 0; i < 128; ++i) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += f(p64[i * 64 + k]) * f(v64[k * 64 + n]); ref2[i * 64 + n] = s; }
  fails += run(2, p64, v64, ref2, 64, "SS B MN-major SW64, N=64 (LBO 4096, SBO 512)");
  run(2, p64, v64, ref2, 64, "  same with LBO / SBO swapped (expected FAIL)", 1);
  {  // P3
    auto p = rnd(128 * 144, 5), v = rnd(144 * 64, 6);
    std::vector<float> ref(128 * 64);
    for (int i = 0; i < 128; ++i) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < 144; ++k) s += f(p[i * 144 + k]) * f(v[k * 64 + n]); ref[i * 64 + n] = s; }
    fails += run(3, p, v, ref, 64, "SS B MN-major SW128, N=64 K=144 (SBO 1024)");
    run(3, p, v, ref, 64, "  same with LBO / SBO swapped (expected FAIL)", 1);
  }
  fails += run(4, p64, v64, ref2, 64, "TS A in TMEM (tcgen05.st bf16 pairs)");
  {
    std::vector<float> ref5(ref2);
    for (int i = 0; i < 64 * 64; ++i) ref5[i] *= 2.0f;
    fails += run(5, p64, v64, ref5, 64, "disable-output-lane mask (lanes 64..127 off)");
  }
  {  // P6
    auto q = rnd(128 * 64, 7), k = rnd(144 * 64, 8);
    std::vector<float> ref(128 * 146);
    for (int i = 0; i < 128; ++i) {
      for (int j = 0; j < 144; ++j) { float s = 0; for (int d = 0; d < 64; ++d) s += f(q[i * 64 + d]) * f(k[j * 64 + d]); ref[i * 146 + j] = s; }
      ref[i * 146 + 144] = ref[i * 146 + 49];
      ref[i * 146 + 145] = ref[i * 146 + 131];
    }
    fails += run(6, q, k, ref, 146, "N=144 accumulator; ld x16 / x1 at odd columns");
  }
  printf("%d probe(s) failed\n", fails);
  return 0;
}
