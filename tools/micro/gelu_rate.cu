// Pure-ALU ceiling of the epilogue GELU variants: elements / clk / SM with W warps per SM, no memory traffic.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../medical_vision_langauge_transformer_b200/csrc/common.cuh"
int mvlt_pdl_enabled(void) { return 0; }
using namespace mvlt;
template <int V> __global__ void k(float* out, int iters, float seed) {
  float2 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = make_float2(seed + i * 0.01f + threadIdx.x * 1e-4f, seed - i * 0.02f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (V == 0) v[i] = gelu_erf_pk2(v[i]);
      else if (V == 1) v[i] = make_float2(gelu_erf_imm(v[i].x), gelu_erf_imm(v[i].y));
      else if (V == 2) v[i] = gelu_erf_fast2(v[i]);
      else { v[i].x = fmaf(v[i].x, 1.0001f, 0.5f); v[i].y = fmaf(v[i].y, 0.9999f, -0.5f); }   // 1 FFMA per element
      v[i].x += 0.37f; v[i].y -= 0.11f;   // keep values in range
    }
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
  if (s == 12345.678f) out[0] = s;
}
template <int V> void run(const char* name, int warps) {
  float* d; cudaMalloc(&d, 4);
  int iters = 2000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<V><<<148, warps * 32>>>(d, 10, 0.3f); cudaDeviceSynchronize();
  cudaEventRecord(a); k<V><<<148, warps * 32>>>(d, iters, 0.3f); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  double elems = 148.0 * warps * 32 * 32 * iters, clk = ms * 1e-3 * 1.965e9;
  printf("%-28s warps/SM %2d: %.2f elem/clk/SM\n", name, warps, elems / clk / 148.0);
}
int main() {
  for (int w : {8, 16, 32}) { run<0>("gelu_erf_pk2 (packed Horner)", w); run<1>("gelu_erf_imm (scalar imm)", w); run<2>("gelu_erf_fast2 (even/odd)", w); run<3>("1 FFMA + 1 FADD / element", w); }
  return 0;
}
