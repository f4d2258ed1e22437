// Micro-benchmark: dense tcgen05.mma kind::tf32 rate on sm_100a (one CTA per SM, one issuing thread, operands
// resident in shared memory, accumulator in TMEM) — the pipe peak the roofline of gram_tc_kernel is judged against.
// Shapes: M = 128, K = 8 per instruction, N = 256 (the widest) and N = 208 (what the k = 100 Gram issues).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I you_can_not_recommend_b200/csrc -I include \
//             -o scripts/micro/tf32_mma_rate scripts/micro/tf32_mma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include "gram_tc.cuh"

using namespace ycnr;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, float* sink, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 8 * 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8 * 4096 / 4; i += 128) reinterpret_cast<float*>(base)[i] = 1.0f + (float)(i % 97) * 0.013f;
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init(smem_u32(bar + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 0) {
    const uint64_t desc0 = tc_smem_desc(smem_u32(base), kTcPanelBytes, 512u, 1u);
    const uint32_t hi = (uint32_t)(desc0 >> 32), lo = (uint32_t)desc0;
    for (int it = 0; it < iters; ++it) {
      asm volatile(
          "{\n .reg .pred q, t;\n .reg .b32 l1, l2, l3;\n .reg .b64 d0, d1, d2, d3;\n"
          " setp.eq.b32 t, 0, 0;\n"
          " add.u32 l1, %1, 64;\n add.u32 l2, l1, 64;\n add.u32 l3, l2, 64;\n"
          " mov.b64 d0, {%1, %2};\n mov.b64 d1, {l1, %2};\n mov.b64 d2, {l2, %2};\n mov.b64 d3, {l3, %2};\n"
          " elect.sync _|q, 0xffffffff;\n"
          " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d0, d0, %3, t;\n"
          " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d1, d1, %3, t;\n"
          " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d2, d2, %3, t;\n"
          " @q tcgen05.mma.cta_group::1.kind::tf32 [%0], d3, d3, %3, t;\n}\n" ::"r"(tmem + (uint32_t)(((mode & 1) ? 0 : (it & 1)) * 256)),
          "r"(lo), "r"(hi), "r"(IDESC)
          : "memory");
      if (mode & 2)   // a commit per group of four MMAs onto a second mbarrier nobody waits for (the Gram kernel's per-stage commit)
        asm volatile(
            "{\n .reg .pred q;\n elect.sync _|q, 0xffffffff;\n"
            " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar + 1))
            : "memory");
    }
    asm volatile(
        "{\n .reg .pred q;\n elect.sync _|q, 0xffffffff;\n"
        " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar))
        : "memory");
    mbar_wait(smem_u32(bar), 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
  if (tid == 0 && sink) sink[blockIdx.x] = 1.f;
}

template <int N>
double run(int sms, int iters, int mode = 0) {
  const size_t smem = 8 * 4096 + 64 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    mma_rate_kernel<N><<<sms, 128, smem>>>(iters, nullptr, mode);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const double flop = 2.0 * 128 * N * 8 * 4.0 * iters * sms;
  return flop / (ms * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, iters = 1 << 16;
  const double t256 = run<256>(sms, iters), t208 = run<208>(sms, iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"tf32_mma_tflops_n256\": %.1f, \"tf32_mma_tflops_n208\": %.1f, \"sms\": %d, \"status\": \"%s\"}\n", t256, t208, sms,
         cudaGetErrorString(e));
  // clocks per MMA instruction (at the nominal 1.965 GHz) for the shapes and issue patterns of the Gram kernel:
  // mode bit 0: ONE accumulator (every MMA depends on the previous one), bit 1: a tcgen05.commit per four MMAs
  for (int mode = 0; mode < 4; ++mode) {
    const double r208 = run<208>(sms, iters, mode), r144 = run<144>(sms, iters, mode), r80 = run<80>(sms, iters, mode), r48 = run<48>(sms, iters, mode);
    auto clk = [&](double tflops, int n) { return 2.0 * 128 * n * 8 / (tflops * 1e12 / sms / 1.965e9); };
    printf("{\"mode\": %d, \"clk_per_mma\": {\"n208\": %.1f, \"n144\": %.1f, \"n80\": %.1f, \"n48\": %.1f}}\n", mode, clk(r208, 208),
           clk(r144, 144), clk(r80, 80), clk(r48, 48));
  }
  return 0;
}
