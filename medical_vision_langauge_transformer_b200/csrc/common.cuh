// Shared device helpers for the MVLT B200 (sm_100a) kernels: vector ld/st, bf16 packing, warp
// reductions, and thin inline-PTX wrappers for mbarrier / TMA / tcgen05 / TMEM.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define MVLT_OK 0
#define MVLT_ERR_INVALID (-1)
#define MVLT_ERR_UNSUPPORTED (-2)
#define MVLT_ERR_DRIVER (-3)

// dtype codes used across the C ABI
#define MVLT_F32 0
#define MVLT_BF16 1

#define MVLT_LAUNCH_CHECK()                         \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)

// Programmatic dependent launch (PDL).  Every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (unless MVLT_PDL=0) and begins with pdl_grid_sync(): the grid may be
// scheduled while its predecessor in the stream drains (its CTAs take SMs as they free up, set up barriers / TMEM /
// descriptors), and only touches global memory after griddepcontrol.wait — which returns once the predecessor has
// completed and flushed.  The dependents of THIS grid are released right after the wait, so look-ahead is one kernel.
int mvlt_pdl_enabled(void);  // c_abi.cu

namespace mvlt {

typedef __nv_bfloat16 bf16;

// Function attributes (opt-in shared memory sizes, carve-outs) are per DEVICE: true the first time the calling thread's
// current device is seen for this `mask` (one mask per attribute group), so a process that drives several GPUs sets them
// on each.  Not synchronised: the reference's callers are single-threaded.
static inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mvlt_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// erf-GELU for the tensor-core epilogues: 9 FP ops + ONE MUFU per element instead of erff's ~36 instructions.
// erfc(|x|/sqrt2) = 2^-q(|x|), q = a*(c1 + a*(c2 + a*(c3 + a*(c4 + a*c5)))): -log2(erfc(t)) fitted on t in [0, 5.5] by
// least squares on the erf error, with the 1/sqrt2 argument scale folded into the coefficients (q grows ~a^2 beyond
// the fit range, so the tail saturates correctly).  Max |erf error| 7.9e-7, max |GELU error| 1.6e-6 over [-12, 12]
// in fp32 arithmetic (fit + check: DESIGN.md §GELU).  The fp32 parity kernels keep erff.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float a = fabsf(x);
  float q = fmaf(a, 0.0005292023415677249f, -0.007443261332809925f);
  q = fmaf(q, a, 0.05264018476009369f);
  q = fmaf(q, a, 0.45920330286026f);
  q = fmaf(q, a, 1.1511013507843018f);
  q = -q * a;
  float erfc_a;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(erfc_a) : "f"(q));
  const float e = copysignf(1.0f - erfc_a, x);
  const float h = 0.5f * x;
  return fmaf(h, e, h);
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 — two fp32 lanes per issued instruction) -------------
__device__ __forceinline__ uint64_t f2_as_u64(float2 v) { return *reinterpret_cast<uint64_t*>(&v); }
__device__ __forceinline__ float2 u64_as_f2(uint64_t v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
  return u64_as_f2(d);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)), "l"(f2_as_u64(c)));
  return u64_as_f2(d);
}

// gelu_erf_fast on a pair, arranged for the packed pipes.  With a = |x|, s = x*x and the same fitted q(a) = c1 a + c2 a^2 +
// c3 a^3 + c4 a^4 + c5 a^5 split into even/odd powers:  q = a*(c1 + s*(c3 + s*c5)) + s*(c2 + s*c4), and
//   GELU(x) = x*Phi(x) = relu(x) - a * 2^(-q - 1)            (Phi(-a) = erfc(a/sqrt2)/2 = 2^(-q-1))
// -> 5 packed FMA/MUL + per lane {FFMA, MUFU.EX2, FMNMX, FFMA}: 6.5 issued instructions per element instead of 11.
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  const float c1 = 1.1511013507843018f, c2 = 0.45920330286026f, c3 = 0.05264018476009369f, c4 = -0.007443261332809925f,
              c5 = 0.0005292023415677249f;
  const float2 s = mul2(x, x);
  float2 A = fma2(s, make_float2(c5, c5), make_float2(c3, c3));
  A = fma2(A, s, make_float2(c1, c1));
  const float2 Bq = fma2(s, make_float2(c4, c4), make_float2(c2, c2));
  const float2 u = fma2(s, Bq, make_float2(1.0f, 1.0f));
  float t0, t1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(-fabsf(x.x), A.x, -u.x)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(-fabsf(x.y), A.y, -u.y)));
  return make_float2(fmaf(-fabsf(x.x), t0, fmaxf(x.x, 0.0f)), fmaf(-fabsf(x.y), t1, fmaxf(x.y, 0.0f)));
}

// Same fit as gelu_erf_fast, arranged for the fma pipe's issue rate (tools/gemm_trace.py: the GELU epilogues are
// fma-pipe bound, ~750 cycles per 32x32 chunk per scheduler; a 3-register FFMA or a packed FFMA2 costs 2 / 4 pipe cycles,
// an FFMA with an immediate operand 1).  GELU(x) = relu(x) - |x| * 2^p(|x|),  p(a) = -1 - q(a) in Horner form with
// literal coefficients (-1: the 1/2 of Phi folded into the exponent): 5 immediate-form FFMA + MUFU.EX2 + FMNMX + 1 FFMA.
__device__ __forceinline__ float gelu_erf_imm(float x) {
  const float a = fabsf(x);
  float p = fmaf(a, -0.0005292023415677249f, 0.007443261332809925f);
  p = fmaf(p, a, -0.05264018476009369f);
  p = fmaf(p, a, -0.45920330286026f);
  p = fmaf(p, a, -1.1511013507843018f);
  p = fmaf(p, a, -1.0f);
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(p));
  return fmaf(-a, t, fmaxf(x, 0.0f));
}

// The same fit once more, for an ISSUE-bound consumer (the GEMM epilogues: ~1 instruction per scheduler cycle while
// they run): 12 issued instructions per PAIR.  With n = -|x| (sign bit OR, alu pipe) the exponent polynomial is
// p(n) = -1 + c1 n - c2 n^2 + c3 n^3 - c4 n^4 + c5 n^5 in packed Horner form (5 FFMA2), then 2 MUFU.EX2, 2 FMNMX for
// relu(x) and one packed FFMA2 for relu(x) + n * 2^p.
__device__ __forceinline__ float2 gelu_erf_pk2(float2 x) {
  const float2 n = make_float2(__uint_as_float(__float_as_uint(x.x) | 0x80000000u), __uint_as_float(__float_as_uint(x.y) | 0x80000000u));
  float2 p = fma2(n, make_float2(0.0005292023415677249f, 0.0005292023415677249f), make_float2(0.007443261332809925f, 0.007443261332809925f));
  p = fma2(p, n, make_float2(0.05264018476009369f, 0.05264018476009369f));
  p = fma2(p, n, make_float2(-0.45920330286026f, -0.45920330286026f));
  p = fma2(p, n, make_float2(1.1511013507843018f, 1.1511013507843018f));
  p = fma2(p, n, make_float2(-1.0f, -1.0f));
  float t0, t1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(p.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(p.y));
  return fma2(n, make_float2(t0, t1), make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

// ---- 4-element (one "quad") typed loads/stores: fp32 -> 16 B, bf16 -> 8 B ---------------------
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f32(float* p, float v) { *p = v; }
__device__ __forceinline__ void from_f32(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- shared-memory address / mbarrier ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug must abort the launch (trap -> cudaErrorLaunchFailure),
// never hang the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// One lane of a converged warp (the compiler keeps the guarded operands in uniform registers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}

// ---- TMA ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D tiled load: c0 = innermost (contiguous) coordinate, c1 = row coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 2-D tiled store smem -> global (bulk async-group completion); out-of-bounds rows/columns are clipped by the TMA unit
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// same, but the tile is ADDED to global memory (fp32 reduction performed at the L2)
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups still READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane_base+i), cols [c, c+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// narrower / store forms of the same 32-lane x 32-bit-column access (thread i <-> lane lane_base+i); any start column
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
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
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// non-blocking phase test (polling loops that watch several barriers)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// barrier among a subset of the CTA's warps (id 1..15; all `nthreads` threads of the subset must call it)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// General UMMA shared-memory matrix descriptor (version 1): start address, leading / stride byte offsets, swizzle layout
// (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).  K-major operands: SBO = bytes between 8-row groups (LBO unused inside one swizzle span);
// MN-major operands (rows = K index, the MN extent contiguous): SBO = bytes between 8-row (K) groups, LBO = bytes between
// swizzle-span-wide column blocks along MN.  Checked on hardware by tools/micro/umma_probe.cu.
constexpr uint32_t UMMA_SW128 = 2, UMMA_SW64 = 4;
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// kind::f16 instruction descriptor with the B operand optionally MN-major (bit 16)
__device__ __forceinline__ uint32_t umma_idesc_bf16_ex(uint32_t m, uint32_t n, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem] restricted to the accumulator lanes whose bit in dis[0..3] is CLEAR
__device__ __forceinline__ void umma_bf16_masked(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                                 uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(d0), "r"(d1), "r"(d2), "r"(d3)
               : "memory");
}
// same with the A operand in tensor memory (bf16 pairs, row i in lane i, 8 columns per 16-wide k-step)
__device__ __forceinline__ void umma_bf16_ts_masked(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                                    uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n}\n"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(d0), "r"(d1), "r"(d2), "r"(d3)
               : "memory");
}

// UMMA shared-memory matrix descriptor for a K-major bf16 tile stored as rows of 128 B with the
// 128-byte swizzle (what a TMA box {64 x rows} with CU_TENSOR_MAP_SWIZZLE_128B produces):
// start>>4 | LBO=1 (unused for swizzled K-major) | SBO = 8 rows * 128 B = 1024 B | version 1 | SW128.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n.
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace mvlt
