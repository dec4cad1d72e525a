// Stand-alone probe (TEST TOOL): how long does one SM take to pull a tile's worth of samples (128 hop blocks of
// 640 bytes = 80 KB) from L2 / HBM into shared memory, when all 148 SMs do it at once?
//   mode 0: one 4-D tensor box {32 floats, 5, 128 hop blocks} with SWIZZLE_128B   (what fe_stream.cu issues)
//   mode 1: one 1-D bulk copy of 80 KB
//   mode 2: five 1-D bulk copies of 16 KB
//   mode 3: cp.async (LDGSTS) 16-byte copies by 128 threads
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int HOP = 160, ROWS = 128, BYTES = ROWS * HOP * 4, ITERS = 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) tma_probe_kernel(const __grid_constant__ CUtensorMap map, const float* src, long long span_blocks,
                                                        int mode, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(mode == 3 ? 129 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
  }
  __syncthreads();
  for (int it = 0; it < ITERS; ++it) {
    // every CTA reads its own 128 hop blocks; the window moves with the iteration
    const long long blk0 = (((long long)it * gridDim.x + blockIdx.x) * ROWS) % (span_blocks - ROWS);
    const float* g = src + blk0 * HOP;
    __syncthreads();
    const long long t0 = clock64();
    if (mode != 3) {
      if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BYTES) : "memory");
        if (mode == 0) {
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
              "l"(&map), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"((int)blk0), "r"(0)
              : "memory");
        } else {
          const int pieces = mode == 1 ? 1 : 5;
          for (int p = 0; p < pieces; ++p)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem) + p * (BYTES / pieces)),
                         "l"((const char*)g + p * (BYTES / pieces)), "r"(BYTES / pieces), "r"(smem_u32(&bar))
                         : "memory");
        }
      }
    } else {
      for (int i = tid; i < BYTES / 16; i += 128)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem) + i * 16), "l"((const char*)g + i * 16) : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24) && !ok; ++spin)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(it & 1)
          : "memory");
    const long long t1 = clock64();
    if (tid == 0) cyc[(long long)blockIdx.x * ITERS + it] = ok ? t1 - t0 : -1;
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  encode_tiled_fn enc = (encode_tiled_fn)fp;
  for (int big = 0; big < 2; ++big) {
    // small: 64 MB (L2 resident after the first pass) ; big: 1.5 GB (HBM)
    const long long span_blocks = big ? (1500LL << 20) / (HOP * 4) : (64LL << 20) / (HOP * 4);
    float* d;
    CK(cudaMalloc(&d, span_blocks * HOP * 4));
    CK(cudaMemset(d, 0, span_blocks * HOP * 4));
    long long* dc;
    CK(cudaMalloc(&dc, sizeof(long long) * sms * ITERS));
    CUtensorMap map;
    const cuuint64_t gdim[4] = {32, 5, (cuuint64_t)span_blocks, 1};
    const cuuint64_t gstride[3] = {128, 640, (cuuint64_t)span_blocks * 640};
    const cuuint32_t box[4] = {32, 5, ROWS, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES + 1024));
    for (int grid : {1, sms}) {
      for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
          tma_probe_kernel<<<grid, 128, BYTES + 1024>>>(map, d, span_blocks, mode, dc);
          CK(cudaDeviceSynchronize());
        }
        std::vector<long long> c((size_t)grid * ITERS);
        CK(cudaMemcpy(c.data(), dc, c.size() * 8, cudaMemcpyDeviceToHost));
        std::vector<long long> steady;
        for (int b = 0; b < grid; ++b)
          for (int it = 4; it < ITERS; ++it) steady.push_back(c[(size_t)b * ITERS + it]);
        std::sort(steady.begin(), steady.end());
        printf("%s grid=%3d mode=%d : cycles per 80 KB load  min %lld  median %lld  p90 %lld  max %lld  (%.1f B/clk median)\n",
               big ? "HBM(1.5GB)" : "L2 (64MB) ", grid, mode, steady.front(), steady[steady.size() / 2], steady[steady.size() * 9 / 10],
               steady.back(), (double)BYTES / steady[steady.size() / 2]);
      }
    }
    cudaFree(d);
    cudaFree(dc);
  }
  printf("probe done\n");
  return 0;
}
