// Micro-benchmark: FP32 FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a), bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, const float* in, int iters) {
  float a[4], b[4];
  for (int i = 0; i < 4; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[1024 + threadIdx.x + 32 * i]; }
  if (MODE == 0) {           // 4x4 outer product, 16 FFMA per step (3 distinct registers each)
    float acc[4][4] = {};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      a[0] += 1e-9f;
    }
    float s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {                   // 16 FFMA2 per step = 32 FMAs (pairs along the reduction index)
    float2 acc[4][4] = {};
    float2 a2[4], b2[4];
    for (int i = 0; i < 4; ++i) { a2[i] = make_float2(a[i], a[3 - i]); b2[i] = make_float2(b[i], b[3 - i]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ffma2(acc[i][j], a2[i], b2[j]);
      a2[0].x += 1e-9f;
    }
    float s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}
int main() {
  float *in, *out; cudaMalloc(&in, 1 << 16); cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMemset(in, 0, 1 << 16);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int iters = 1 << 16, grid = p.multiProcessorCount * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, 512>>>(out, in, iters); else k<1><<<grid, 512>>>(out, in, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)grid * 512 * iters * 16 * (mode ? 2 : 1);
    printf("%s: %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM @1.965GHz\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / p.multiProcessorCount / 1.965e9);
  }
  return 0;
}
