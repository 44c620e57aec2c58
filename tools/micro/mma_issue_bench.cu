// Micro-benchmark: how fast can one thread issue tcgen05.mma (bf16, M=128, K=16) instructions that accumulate into the
// same / into alternating TMEM accumulators?  Answers whether short MMAs (N=64/128) are throughput- or latency-paced.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I amodal-depth-anything_b200/csrc -o gpurun_out/mma_issue_bench tools/micro/mma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ada;

template <int N, int NACC, int TS, int WARP = 0>
__global__ void __launch_bounds__(128, 1) k(long long* out, int count) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 65536, tptr = sbase + 65536 + 16;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (WARP && threadIdx.x < 32) {  // whole-warp issue, one elected lane per instruction
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const long long t0 = clock64();
    const uint64_t da0 = make_smem_desc_sw128(sbase, 16, 1024), db0 = make_smem_desc_sw128(sbase + 16384, 16, 1024);
    for (int i = 0; i < count; ++i) {
      const uint32_t d = tmem + (i % NACC) * N;
      if (TS)
        umma_bf16_ts_w(d, tmem + 448 + (i & 3) * 8, db0 + 2 * (i & 3), idesc, 1u);
      else
        umma_bf16_ss_w(d, da0 + 2 * (i & 3), db0 + 2 * (i & 3), idesc, 1u);
    }
    umma_commit_w(bar);
    const long long t1 = clock64();
    mbar_wait(bar, 0, 0x900);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  } else if (!WARP && threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint64_t da = make_smem_desc_sw128(sbase + (i & 3) * 32, 16, 1024);
      const uint64_t db = make_smem_desc_sw128(sbase + 16384 + (i & 3) * 32, 16, 1024);
      const uint32_t d = tmem + (i % NACC) * N;
      if (TS)
        umma_bf16_ts(d, tmem + 448 + (i & 3) * 8, db, idesc, 1u);
      else
        umma_bf16_ss(d, da, db, idesc, 1u);
    }
    umma_commit(bar);
    const long long t1 = clock64();
    mbar_wait(bar, 0, 0x900);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, int NACC, int TS, int WARP = 0>
void run(long long* d, int grid) {
  cudaFuncSetAttribute(k<N, NACC, TS, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  const int count = 256;
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    k<N, NACC, TS, WARP><<<grid, 128, 65536 + 64>>>(d, count);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d nacc=%d: %s\n", N, NACC, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s%s M=128 N=%3d accumulators=%d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (full-rate floor %d)\n", WARP ? "warp-issue " : "", TS ? "TS" : "SS", N,
         NACC, grid, double(h[0]) / count, double(h[1]) / count, N / 2);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int grid : {1, 148}) {
    run<64, 1, 0>(d, grid);  run<64, 2, 0>(d, grid);  run<64, 4, 0>(d, grid);
    run<128, 1, 0>(d, grid); run<128, 2, 0>(d, grid); run<128, 3, 0>(d, grid);
    run<256, 1, 0>(d, grid);
    run<64, 1, 1>(d, grid);  run<64, 2, 1>(d, grid);  run<64, 4, 1>(d, grid);
    run<64, 1, 0, 1>(d, grid); run<128, 1, 0, 1>(d, grid); run<256, 1, 0, 1>(d, grid); run<64, 1, 1, 1>(d, grid);
  }
  return 0;
}
