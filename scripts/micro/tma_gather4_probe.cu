// Probe: what does cp.async.bulk.tensor.2d ... tile::gather4 write to shared memory for a [rows x k] fp32 matrix,
// box = {32 columns, 1 row}, for each swizzle mode?  Prints, per mode, the (row, col) found at every 16-byte chunk of
// the 512-byte destination, so the layout can be compared with the UMMA operand atom of gram_tc.cuh
//   byte = (r % 4) * 128 + ((chunk32 ^ (r % 4)) * 32) + (byte % 32).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/micro/tma_gather4_probe scripts/micro/tma_gather4_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int col, int r0, int r1, int r2, int r3, float* out, int nfloat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  float* dst = reinterpret_cast<float*>(base);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 4096);
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) dst[i] = -1.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(512u) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(b), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
    uint32_t ok = 0;
    for (int spin = 0; spin < 2000000 && !ok; ++spin) {       // bounded: a wrong byte count must not hang the box
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(b), "r"(0u) : "memory");
    }
    if (!ok) dst[255] = -7.f;                                  // marker: the barrier never completed
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = dst[i];
}

int main() {
  const int R = 64, K = 100;
  std::vector<float> h((size_t)R * K);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < K; ++c) h[(size_t)r * K + c] = r * 1000.f + c;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int NF = 256;
  cudaMalloc(&out, NF * 4);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const CUtensorMapSwizzle modes[] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};
  const char* names[] = {"NONE", "128B", "128B_ATOM_32B"};
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  for (int box1 = 1; box1 <= 4; box1 += 3)
  for (int m = 0; m < 3; ++m)
    for (int col : {0, 96}) {
      CUtensorMap tmap;
      cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)R};
      cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
      cuuint32_t box[2] = {32, (cuuint32_t)box1};
      cuuint32_t estr[2] = {1, 1};
      CUresult rc = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[m],
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("== box {32,%d} swizzle %s col %d: encode rc=%d\n", box1, names[m], col, (int)rc);
      if (rc != CUDA_SUCCESS) continue;
      cudaMemset(out, 0, NF * 4);
      probe<<<1, 32, 8192>>>(tmap, col, 5, 17, 3, 60, out, NF);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("   kernel: %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<float> o(NF);
      cudaMemcpy(o.data(), out, NF * 4, cudaMemcpyDeviceToHost);
      if (o[255] == -7.f) printf("   (mbarrier did not complete: byte count != 512)\n");
      for (int chunk = 0; chunk < NF / 4; ++chunk) {
        const float v = o[chunk * 4];
        if (chunk % 8 == 0) printf("   byte %4d:", chunk * 16);
        if (v < 0) printf("   ----- "); else printf(" r%02d.c%03d", (int)(v / 1000), (int)(v - 1000 * (int)(v / 1000)));
        if (chunk % 8 == 7) printf("\n");
      }
    }
  return 0;
}
