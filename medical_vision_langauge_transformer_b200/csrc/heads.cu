// Task-head tail kernels (latency-sized; one warp per row):
//   mvlt_linear_small   nn.Linear with a handful of outputs — retrieval Linear(768,2) model.py:435,
//                       ITM Linear(768,2) model.py:363 (and any N <= 16 head)
//   mvlt_softmax_rows   model.py:348 (VQA softmax over result_num), model.py:468 (retrieval softmax over 2)
//   mvlt_masked_ce_rows F.cross_entropy(ignore_index=-100) over fp32 logits rows, model.py:410,:418
#include "common.cuh"

namespace mvlt {

template <typename TX>
__global__ void __launch_bounds__(256)
linear_small_kernel(const TX* __restrict__ x, long long ldx, const float* __restrict__ w, const float* __restrict__ b,
                    float* __restrict__ out, long long rows, int N, int K) {
  pdl_grid_sync();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TX* xr = x + row * ldx;
  for (int n = 0; n < N; ++n) {
    float acc = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 a = load4(xr + k);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (long long)n * K + k));
      acc = fmaf(a.x, ww.x, acc); acc = fmaf(a.y, ww.y, acc); acc = fmaf(a.z, ww.z, acc); acc = fmaf(a.w, ww.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row * N + n] = acc + (b ? b[n] : 0.f);
  }
}

__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int N) {
  pdl_grid_sync();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* r = in + row * N;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < N; j += 32) sum += expf(r[j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < N; j += 32) out[row * N + j] = expf(r[j] - mx) * inv;
}

// per-row loss = logsumexp(row) - row[label]; rows with label == ignore contribute nothing.
// loss_sum[0] += loss, loss_sum[1] += 1 per counted row (host divides: mean over non-ignored rows).
__global__ void __launch_bounds__(256)
masked_ce_rows_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                      float* __restrict__ loss_sum, long long rows, int N, long long ignore) {
  pdl_grid_sync();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long lab = labels[row];
  if (lab == ignore) return;
  const float* r = logits + row * ld;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < N; j += 32) sum += expf(r[j] - mx);
  sum = warp_sum(sum);
  if (lane == 0) {
    atomicAdd(&loss_sum[0], logf(sum) + mx - r[lab]);
    atomicAdd(&loss_sum[1], 1.0f);
  }
}

}  // namespace mvlt

using namespace mvlt;

extern "C" int mvlt_linear_small(const void* x, int x_dtype, long long ldx, const float* w, const float* bias,
                                 float* out, long long rows, int N, int K, cudaStream_t stream) {
  if (!x || !w || !out || rows <= 0 || N <= 0 || N > 16 || K % 4 || ldx % 4) return MVLT_ERR_INVALID;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (x_dtype == MVLT_F32) launch_k(linear_small_kernel<float>, dim3(grid), dim3(256), 0, stream, (const float*)x, ldx, w, bias, out, rows, N, K);
  else if (x_dtype == MVLT_BF16) launch_k(linear_small_kernel<bf16>, dim3(grid), dim3(256), 0, stream, (const bf16*)x, ldx, w, bias, out, rows, N, K);
  else return MVLT_ERR_INVALID;
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_softmax_rows(const float* in, float* out, long long rows, int N, cudaStream_t stream) {
  if (!in || !out || rows <= 0 || N <= 0) return MVLT_ERR_INVALID;
  launch_k(softmax_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, stream, in, out, rows, N);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_masked_ce_rows(const float* logits, long long ld, const long long* labels, float* loss_sum,
                                   long long rows, int N, long long ignore_index, cudaStream_t stream) {
  if (!logits || !labels || !loss_sum || rows <= 0 || N <= 0) return MVLT_ERR_INVALID;
  cudaError_t e = cudaMemsetAsync(loss_sum, 0, 2 * sizeof(float), stream);
  if (e != cudaSuccess) return (int)e;
  launch_k(masked_ce_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, stream, logits, ld, labels, loss_sum, rows, N, ignore_index);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}
