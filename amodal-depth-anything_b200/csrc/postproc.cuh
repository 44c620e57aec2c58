// Device-side pre/post-processing of the reference's single-image inference (infer.py:16-28, 72-103; SURVEY.md section 8
// row f2). In the reference these steps bounce through the host (depth_raw.cpu() -> normalise -> .cuda(), pred.cpu() ->
// numpy blend -> cv2.blur); here they are small bandwidth kernels on the forward's stream. All fp32, single image.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ada {

// ATen nearest (F.interpolate default / torchvision Resize(NEAREST), infer.py:85-87,97-100): src = min(floor(dst * in/out), in-1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) {
  return min(static_cast<int>(floorf(static_cast<float>(dst) * scale)), in - 1);
}

// uint8 HWC image (cv2 order, infer.py:76) -> fp32 CHW in [0,1] at (H, W) by nearest sampling: rgb/255 then Resize(NEAREST)
// (infer.py:84-86). normalize != 0 additionally applies (x - mean) / std with the ImageNet constants per channel index
// (infer.py:18, the un-guided model's input). One thread per output pixel, 3 channels.
__global__ void __launch_bounds__(256)
image_nearest_kernel(const uint8_t* __restrict__ img, int H0, int W0, float* __restrict__ out, int H, int W, int normalize) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i - y * W;
  const int sy = nearest_src(y, static_cast<float>(H0) / static_cast<float>(H), H0);
  const int sx = nearest_src(x, static_cast<float>(W0) / static_cast<float>(W), W0);
  const uint8_t* p = img + (static_cast<long long>(sy) * W0 + sx) * 3;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = static_cast<float>(p[c]) / 255.0f;
    if (normalize) v = __fdiv_rn(__fsub_rn(v, mean[c]), sd[c]);  // same two roundings as the tensor expression
    out[static_cast<long long>(c) * H * W + i] = v;
  }
}

// uint8 mask (any non-zero = inside, infer.py:80-81) -> nearest resize -> mask01 (0/1, infer.py:87,100-101) and the
// network guide mask01 * 2 - 1 (infer.py:91). Either output may be null.
__global__ void __launch_bounds__(256)
mask_nearest_kernel(const uint8_t* __restrict__ mask, int H0, int W0, float* __restrict__ mask01, float* __restrict__ guide,
                    int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i - y * W;
  const int sy = nearest_src(y, static_cast<float>(H0) / static_cast<float>(H), H0);
  const int sx = nearest_src(x, static_cast<float>(W0) / static_cast<float>(W), W0);
  const float m = mask[static_cast<long long>(sy) * W0 + sx] > 0 ? 1.0f : 0.0f;
  if (mask01) mask01[i] = m;
  if (guide) guide[i] = m * 2.0f - 1.0f;
}

// min / max of a fp32 map (infer.py:22). Order-preserving float <-> uint mapping so atomicMin/Max work for any sign.
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__global__ void minmax_init_kernel(uint32_t* mm) {
  mm[0] = 0xffffffffu;  // min
  mm[1] = 0u;           // max
}
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ d, long long n, uint32_t* mm) {
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t o = f2ord(d[i]);
    lo = min(lo, o);
    hi = max(hi, o);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, lo);
    atomicMax(mm + 1, hi);
  }
}
// base01 = (d - min) / (max - min) (infer.py:22); observation = base01 * 2 - 1 (infer.py:92). Either output may be null.
__global__ void __launch_bounds__(256)
normalize_kernel(const float* __restrict__ d, long long n, const uint32_t* __restrict__ mm, float* __restrict__ base01,
                 float* __restrict__ obs) {
  const float lo = ord2f(mm[0]), hi = ord2f(mm[1]);
  const float range = hi - lo;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = (d[i] - lo) / range;
    if (base01) base01[i] = v;
    if (obs) obs[i] = v * 2.0f - 1.0f;
  }
}

// median_filter_blend (infer.py:30-44, filter_width 3): blended = mask ? amodal : raw; the seam = pixels whose 3x3
// zero-padded mask sum is in (0, 9) is replaced by the 3x3 box mean of `blended` (cv2.blur: BORDER_REFLECT_101, row sums
// then column sum, times 1/9 in fp32).
__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) return -i;
  if (i >= n) return 2 * n - 2 - i;
  return i;
}
__global__ void __launch_bounds__(256)
blend_seam_kernel(const float* __restrict__ raw, const float* __restrict__ amodal, const float* __restrict__ mask01,
                  float* __restrict__ out, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i - y * W;
  auto blended = [&](int yy, int xx) {
    const int j = yy * W + xx;
    return mask01[j] > 0.f ? amodal[j] : raw[j];
  };
  float dil = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) dil += mask01[yy * W + xx];
    }
  float v = blended(y, x);
  if (dil > 0.f && dil < 9.f) {
    float rows[3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = reflect101(y + dy, H);
      rows[dy + 1] = (blended(yy, reflect101(x - 1, W)) + blended(yy, x)) + blended(yy, reflect101(x + 1, W));
    }
    v = ((rows[0] + rows[1]) + rows[2]) * (1.0f / 9.0f);
  }
  out[i] = v;
}

}  // namespace ada
