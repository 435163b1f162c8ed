// Task-head tail kernels (latency-sized; one warp per row):
//   mvlt_linear_small   nn.Linear with a handful of outputs — retrieval Linear(768,2) model.py:435,
//                       ITM Linear(768,2) model.py:363 (and any N <= 16 head)
//   mvlt_softmax_rows   model.py:348 (VQA softmax over result_num), model.py:468 (retrieval softmax over 2)
//   mvlt_masked_ce_rows F.cross_entropy(ignore_index=-100) over fp32 logits rows, model.py:410,:418
//   mvlt_rank_first_positive  run_retrieval.py:220-249 `compute_ranks` on the device: per image row (and per caption
//                       column) the position of the best-placed positive in descending-score order
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
  if (lab < 0 || lab >= N) {      // F.cross_entropy raises here; a kernel cannot: poison the loss instead of reading out of bounds
    if (lane == 0) atomicAdd(&loss_sum[0], __int_as_float(0x7fc00000));
    return;
  }
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

// Finishing pass of the fused vocabulary projection + cross-entropy (model.py:404-410): per labelled row
//   loss = logsumexp(logits) - logits[label],  logsumexp from the GEMM epilogue's (max, sum exp) partials,
//   logits[label] = t[row,:] . W[label,:] + bias[label] recomputed from the bf16 operands (fp32 accumulate, as the GEMM does).
// Rows with label == ignore contribute nothing; a label outside [0, N) poisons the loss with NaN (F.cross_entropy raises).
__global__ void __launch_bounds__(256)
mlm_ce_rows_kernel(const float2* __restrict__ part, long long ld_part, int n_part, const bf16* __restrict__ t, long long ldt,
                   const bf16* __restrict__ w, long long ldw, const float* __restrict__ bias, const long long* __restrict__ labels,
                   float* __restrict__ row_loss, long long rows, int N, int K, long long ignore) {
  pdl_grid_sync();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long lab = labels[row];
  if (lab == ignore) { if (lane == 0) row_loss[row] = 0.f; return; }
  if (lab < 0 || lab >= N) { if (lane == 0) row_loss[row] = __int_as_float(0x7fc00000); return; }
  const float2* pr = part + row * ld_part;
  float mx = -INFINITY;
  for (int j = lane; j < n_part; j += 32) mx = fmaxf(mx, pr[j].x);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n_part; j += 32) sum += pr[j].y * expf(pr[j].x - mx);
  sum = warp_sum(sum);
  float dot = 0.f;
  const bf16* tr = t + row * ldt;
  const bf16* wr = w + lab * ldw;
  for (int k = lane * 2; k < K; k += 64) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(tr + k)), b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(wr + k));
    dot = fmaf(a.x, b.x, dot);
    dot = fmaf(a.y, b.y, dot);
  }
  dot = warp_sum(dot);
  if (lane == 0) row_loss[row] = logf(sum) + mx - (dot + (bias ? bias[lab] : 0.f));
}

// deterministic reduction of the per-row losses: one CTA, fixed order -> (sum over labelled rows, number of labelled rows)
__global__ void __launch_bounds__(256)
ce_reduce_kernel(const float* __restrict__ row_loss, const long long* __restrict__ labels, float* __restrict__ loss_sum, long long rows,
                 long long ignore) {
  pdl_grid_sync();
  __shared__ float s_l[256], s_c[256];
  float l = 0.f, c = 0.f;
  for (long long i = threadIdx.x; i < rows; i += 256)
    if (labels[i] != ignore) { l += row_loss[i]; c += 1.f; }
  s_l[threadIdx.x] = l; s_c[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s_l[threadIdx.x] += s_l[threadIdx.x + o]; s_c[threadIdx.x] += s_c[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { loss_sum[0] = s_l[0]; loss_sum[1] = s_c[0]; }
}

// Position of the first positive of one line of the score matrix in descending order.  np.argsort(sim)[::-1] lists equal
// scores by DEcreasing index (stable ascending sort, reversed), so the order is the lexicographic key (score, index)
// descending: the first positive is the positive with the largest key and its rank is the number of elements with a
// larger key — two O(n) passes, no sort.  LINE_IS_ROW: one warp per row (coalesced along the row); otherwise one thread
// per column (adjacent threads read adjacent columns).
__device__ __forceinline__ bool key_gt(float s, int j, float s0, int j0) { return s > s0 || (s == s0 && j > j0); }

__global__ void __launch_bounds__(256)
rank_rows_kernel(const float* __restrict__ scores, long long lds, const unsigned char* __restrict__ labels, long long ldl,
                 int* __restrict__ out, int R, int C, int none) {
  pdl_grid_sync();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* s = scores + (long long)row * lds;
  const unsigned char* l = labels + (long long)row * ldl;
  float bs = -INFINITY;
  int bj = -1;
  for (int j = lane; j < C; j += 32)
    if (l[j] == 1 && (bj < 0 || key_gt(s[j], j, bs, bj))) { bs = s[j]; bj = j; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (oj >= 0 && (bj < 0 || key_gt(os, oj, bs, bj))) { bs = os; bj = oj; }
  }
  int cnt = 0;
  if (bj >= 0)
    for (int j = lane; j < C; j += 32) cnt += key_gt(s[j], j, bs, bj) ? 1 : 0;
  cnt = (int)warp_sum((float)cnt);   // exact for counts < 2^24
  if (lane == 0) out[row] = bj >= 0 ? cnt : none;
}

__global__ void __launch_bounds__(128)
rank_cols_kernel(const float* __restrict__ scores, long long lds, const unsigned char* __restrict__ labels, long long ldl,
                 int* __restrict__ out, int R, int C, int none) {
  pdl_grid_sync();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= C) return;
  float bs = -INFINITY;
  int bi = -1;
  for (int i = 0; i < R; ++i) {
    const float v = scores[(long long)i * lds + col];
    if (labels[(long long)i * ldl + col] == 1 && (bi < 0 || key_gt(v, i, bs, bi))) { bs = v; bi = i; }
  }
  int cnt = 0;
  if (bi >= 0)
    for (int i = 0; i < R; ++i) cnt += key_gt(scores[(long long)i * lds + col], i, bs, bi) ? 1 : 0;
  out[col] = bi >= 0 ? cnt : none;
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

extern "C" int mvlt_gemm_bf16_lse_partials(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* part,
                                           long long ld_part, int M, int N, int K, cudaStream_t stream);

extern "C" long long mvlt_mlm_ce_workspace_bytes(long long rows, int N) {
  const long long tiles_n = (N + 255) / 256;
  return rows * 2 * tiles_n * 8 + ((rows * 4 + 15) / 16) * 16;      // (max, sum) partials + per-row losses
}

// Masked-LM loss without the logits: t bf16 [rows, K] (transformed + normalised text rows), w bf16 [N, K] (decoder weight),
// bias fp32 [N], labels int64 [rows] -> loss_sum[0] = sum of per-row cross-entropies over rows with label != ignore_index,
// loss_sum[1] = their count.  workspace: mvlt_mlm_ce_workspace_bytes(rows, N) bytes, 16-byte aligned.  Replaces the decoder
// nn.Linear(768, vocab) of HF modeling_bert.py:502-512 + nn.CrossEntropyLoss(ignore_index=-100) of model.py:404-410.
extern "C" int mvlt_mlm_ce_fused(const void* t, long long ldt, const void* w, long long ldw, const float* bias, const long long* labels,
                                 float* loss_sum, void* workspace, long long workspace_bytes, long long rows, int N, int K,
                                 long long ignore_index, cudaStream_t stream) {
  if (!t || !w || !labels || !loss_sum || !workspace || rows <= 0 || N <= 0 || K <= 0 || K % 2) return MVLT_ERR_INVALID;
  if (rows > 0x7fffffffLL || workspace_bytes < mvlt_mlm_ce_workspace_bytes(rows, N) || ((uintptr_t)workspace & 15)) return MVLT_ERR_INVALID;
  const long long tiles_n = (N + 255) / 256, ld_part = 2 * tiles_n;
  float* row_loss = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + rows * ld_part * 8);
  int rc = mvlt_gemm_bf16_lse_partials(t, ldt, w, ldw, bias, workspace, ld_part, (int)rows, N, K, stream);
  if (rc != MVLT_OK) return rc;
  launch_k(mlm_ce_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, stream, reinterpret_cast<const float2*>(workspace), ld_part,
           (int)ld_part, reinterpret_cast<const bf16*>(t), ldt, reinterpret_cast<const bf16*>(w), ldw, bias, labels, row_loss, rows, N, K,
           ignore_index);
  MVLT_LAUNCH_CHECK();
  launch_k(ce_reduce_kernel, dim3(1), dim3(256), 0, stream, (const float*)row_loss, labels, loss_sum, rows, ignore_index);
  MVLT_LAUNCH_CHECK();
  return MVLT_OK;
}

extern "C" int mvlt_rank_first_positive(const float* scores, long long lds, const unsigned char* labels, long long ldl,
                                        int* row_ranks, int* col_ranks, int R, int C, cudaStream_t stream) {
  if (!scores || !labels || R <= 0 || C <= 0 || lds < C || ldl < C || (!row_ranks && !col_ranks)) return MVLT_ERR_INVALID;
  if (R >= (1 << 24) || C >= (1 << 24)) return MVLT_ERR_UNSUPPORTED;
  if (row_ranks) {   // a row without a positive ranks `C`, a column without one `C` as well (run_retrieval.py:230,:242)
    mvlt::launch_k(mvlt::rank_rows_kernel, dim3((R + 7) / 8), dim3(256), 0, stream, scores, lds, labels, ldl, row_ranks, R, C, C);
    MVLT_LAUNCH_CHECK();
  }
  if (col_ranks) {
    mvlt::launch_k(mvlt::rank_cols_kernel, dim3((C + 127) / 128), dim3(128), 0, stream, scores, lds, labels, ldl, col_ranks, R, C, C);
    MVLT_LAUNCH_CHECK();
  }
  return MVLT_OK;
}
