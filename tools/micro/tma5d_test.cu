// Micro-test: 5-D TMA store (pixel-shuffle view of a transposed-conv output) with in-bounds, negative and overflowing x.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I amodal-depth-anything_b200/csrc tools/micro/tma5d_test.cu -o /tmp/tma5d && /tmp/tma5d
#include <cstdio>
#include <vector>
#include "ptx.cuh"
#include "tma_host.h"
using namespace ada;

__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int yb, int ky, int kx, int co0) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = smem_u32(sm);
  // 32 rows x 64 bf16, value = row (no swizzle correctness needed here: every element of a row holds the row index)
  for (int i = threadIdx.x; i < 32 * 64; i += 32) reinterpret_cast<__nv_bfloat16*>(sm)[i] = __float2bfloat16(float(i / 64 + 1));
  fence_proxy_async_smem();
  __syncwarp();
  tma_store_5d_w(&tm, base, co0, kx, x, ky, yb);
  bulk_commit_w();
  bulk_wait_w<0>();
}

int main(int argc, char** argv) {
  const int Cout = 256, ks = 4, W = 37, BH = 74;
  const size_t n = (size_t)BH * ks * W * ks * Cout;
  __nv_bfloat16* out;
  cudaMalloc(&out, n * 2);
  uint64_t dims[5] = {Cout, ks, W, ks, BH};
  uint64_t str[4] = {Cout * 2ull, (uint64_t)ks * Cout * 2, (uint64_t)W * ks * Cout * 2, (uint64_t)ks * W * ks * Cout * 2};
  uint32_t box[5] = {64, 1, 32, 1, 1};
  CUtensorMap tm = make_tmap_bf16(out, 5, dims, str, box);
  struct T { int x, yb; const char* what; } tests[] = {{0, 0, "in bounds x=0"}, {5, 3, "overflow x (5+32>37)"}, {-7, 4, "negative x"}, {0, 73, "last yb"}, {3, 80, "yb out of range"}};
  for (auto& t : tests) {
    cudaMemset(out, 0, n * 2);
    k<<<1, 32, 4096>>>(tm, t.x, t.yb, 2, 1, 64);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> h(n);
    cudaMemcpy(h.data(), out, n * 2, cudaMemcpyDeviceToHost);
    size_t nz = 0;
    for (auto v : h) nz += v != 0;
    printf("%-24s -> %s, nonzero elements %zu (expect %d)\n", t.what, cudaGetErrorString(e), nz,
           64 * (t.yb >= BH ? 0 : (t.x < 0 ? 32 + t.x : (t.x + 32 > W ? W - t.x : 32))));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
