// fp32 "parity mode" GEMM on the CUDA cores (FFMA, fp32 accumulate, no tensor-core rounding):
//     C[M,N] = epilogue( A[M,K] . W[N,K]^T )
// Same call sites and epilogues as gemm_tc.cu; used when the model runs with precision="fp32" so the
// forward can be checked against the reference at 1e-4 relative (BASELINE.json north_star).  Register-tiled
// 64x64x16, 256 threads, 4x4 outputs per thread, smem tiles stored K-major-transposed for conflict-free reads.
#include "common.cuh"

namespace mvlt {

constexpr int SBM = 64, SBN = 64, SBK = 16;

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                 float* __restrict__ C, long long ldc, const float* __restrict__ bias,
                 const float* __restrict__ res, long long ldres, int M, int N, int K, int act) {
  pdl_grid_sync();
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int lr = tid >> 2;        // tile row loaded by this thread (0..63)
  const int lk = (tid & 3) * 4;   // k offset of its float4
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SBK) {
    float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
    if (m0 + lr < M && k0 + lk < K) a = load4(A + (long long)(m0 + lr) * lda + k0 + lk);
    if (n0 + lr < N && k0 + lk < K) b = load4(W + (long long)(n0 + lr) * ldw + k0 + lk);
    As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
    Bs[lk + 0][lr] = b.x; Bs[lk + 1][lr] = b.y; Bs[lk + 2][lr] = b.z; Bs[lk + 3][lr] = b.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (act == 1) v = gelu_erf(v);
      else if (act == 2) v = tanhf(v);
      if (res) v += res[(long long)m * ldres + n];
      if (act >= 3) {  // ResNet epilogues, applied after the residual add: 3 ReLU, 4 ReLU then erf-GELU
        v = fmaxf(v, 0.f);
        if (act == 4) v = gelu_erf(v);
      }
      C[(long long)m * ldc + n] = v;
    }
  }
}

}  // namespace mvlt

extern "C" int mvlt_gemm_f32_simt(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc,
                                  const float* bias, const float* residual, long long ldres, int M, int N, int K,
                                  int act, cudaStream_t stream) {
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || K % 4 != 0 || lda % 4 != 0 || ldw % 4 != 0) return MVLT_ERR_INVALID;
  if (((uintptr_t)A & 15) || ((uintptr_t)W & 15)) return MVLT_ERR_INVALID;
  dim3 grid((N + mvlt::SBN - 1) / mvlt::SBN, (M + mvlt::SBM - 1) / mvlt::SBM);
  if (grid.y > 65535) return MVLT_ERR_UNSUPPORTED;
  mvlt::launch_k(mvlt::gemm_simt_kernel, dim3(grid), dim3(256), 0, stream, A, lda, W, ldw, C, ldc, bias, residual, ldres, M, N, K, act);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
