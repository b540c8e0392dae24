// tools/tma_stream.cu -- micro-benchmark: how fast can ONE SM stream L2-resident weights through a TMA ring?
// Answers what bounds the fused FFN kernel (mlp_tc.cu): ring depth x latency, or a per-SM / aggregate bandwidth limit.
// Build + run (on the GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Imemotr_b200/csrc -Iinclude \
//                                  tools/tma_stream.cu -o gpurun_out/tma_stream -lcuda && gpurun_out/tma_stream
#include <cstdio>
#include <vector>
#include "tc_common.cuh"

namespace memotr { thread_local char g_err[512]; }   // common.cuh's error buffer lives in capi.cu in the library
using namespace memotr;
using namespace memotr::tc;

constexpr int SLOT = 32768;

__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap tmW, int nslot, int iters, int rows_total, long long *clk_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + nslot * SLOT), *empty = full + 8;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslot; ++s) mbar_init(full + s, 1), mbar_init(empty + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int t = 0; t < iters; ++t) {
      const int s = t % nslot;
      mbar_wait(empty + s, ((t / nslot) & 1) ^ 1);
      mbar_expect_tx(full + s, SLOT);
      const int r = ((t + blockIdx.x * 7) * 128) % rows_total;      // 128 rows x 256 cols bf16 = 2 boxes of 16 KB
      tma_load_2d(smem + s * SLOT, &tmW, full + s, 0, r);
      tma_load_2d(smem + s * SLOT + 16384, &tmW, full + s, 64, r);
    }
  } else if (threadIdx.x == 32) {
    for (int t = 0; t < iters; ++t) {
      const int s = t % nslot;
      mbar_wait(full + s, (t / nslot) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clk_out[blockIdx.x] = clock64() - t0;
}

int main() {
  const int rows = 4096, cols = 256;                 // 2 MB of bf16 "weights": L2-resident
  void *W;
  cudaMalloc(&W, (size_t)rows * cols * 2);
  cudaMemset(W, 0, (size_t)rows * cols * 2);
  long long *clk;
  cudaMalloc(&clk, 1024 * sizeof(long long));
  CUtensorMap tm;
  if (!make_map(&tm, W, rows, cols, cols, 128)) { printf("make_map failed\n"); return 1; }
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * SLOT + 2048);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int iters = 2048;                            // 64 MB per CTA
  printf("nslot  ctas  in-flight/SM  ms      GB/s/SM   TB/s total  B/clk/SM(clock64)\n");
  for (int ctas : {1, 16, 74, 148, 296}) {
    for (int nslot : {1, 2, 3, 4, 6}) {
      if (ctas == 296 && nslot > 3) continue;        // two CTAs per SM need <= 113 KB each
      const size_t sm = (size_t)nslot * SLOT + 2048;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<<<ctas, 64, sm>>>(tm, nslot, iters, rows, clk);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      if (cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> h(ctas);
      cudaMemcpy(h.data(), clk, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (auto v : h) mx = v > mx ? v : mx;
      const double bytes = (double)iters * SLOT;
      const int per_sm = ctas > 148 ? 2 : 1;
      printf("%5d %5d %9d KB  %7.3f  %7.1f  %9.2f   %7.1f\n", nslot, ctas, nslot * 32 * per_sm, ms,
             bytes * per_sm / ms / 1e6, bytes * ctas / ms / 1e9, bytes / (double)mx);
    }
  }
  return 0;
}
