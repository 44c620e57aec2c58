// Per-sample evaluation post-ops of the reference's validation loop on the device (discriminative_trainer.py:542-613;
// SURVEY.md section 8 row f3): nearest resize of the prediction to the ground-truth size (:542), least-squares scale/shift
// alignment against the observation over the visible mask (src/util/alignment.py:7-54, solved as 2x2 normal equations
// instead of np.linalg.lstsq) and the ten masked metrics of src/util/metric.py:37-161 for the raw and the aligned
// prediction (:584-613). The reference does this on the host with a .cpu().numpy() round trip per sample.
// Per-pixel terms are evaluated in fp32 like the torch expressions; sums are accumulated in fp64.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ada {

constexpr int kEvalAlignSums = 5;    // n, sum p, sum p^2, sum g, sum p g          (visible mask)
constexpr int kEvalMetricSums = 10;  // absrel, sqrel, se, sle, sl, l10, d1, d2, d3, inv-se (object mask), x2 variants
constexpr int kEvalScratch = kEvalAlignSums + 1 + 2 * kEvalMetricSums;  // doubles
constexpr int kEvalOut = 24;  // scale, shift, 10 metrics (pred), 10 metrics (aligned), n_visible, n_object

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

struct EvalArgs {
  const float* pred;  // [h, w] network output
  int h, w;
  const float* gt;    // [H, W] depth_raw_ts
  const float* obs;   // [H, W] depth_observation
  const uint8_t* visible;  // [H, W]
  const uint8_t* object;   // [H, W]
  int H, W;
  double* scratch;    // kEvalScratch doubles, zeroed by the launcher
  double* out;        // kEvalOut doubles
};

__device__ __forceinline__ float eval_pred_at(const EvalArgs& a, int i) {
  const int y = i / a.W, x = i - y * a.W;
  const int sy = min(static_cast<int>(floorf(static_cast<float>(y) * (static_cast<float>(a.h) / static_cast<float>(a.H)))), a.h - 1);
  const int sx = min(static_cast<int>(floorf(static_cast<float>(x) * (static_cast<float>(a.w) / static_cast<float>(a.W)))), a.w - 1);
  return a.pred[sy * a.w + sx];
}

__global__ void __launch_bounds__(256) eval_align_sums_kernel(const EvalArgs a) {
  double s[kEvalAlignSums] = {0, 0, 0, 0, 0};
  const int n = a.H * a.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (a.visible[i]) {
      const double p = eval_pred_at(a, i), g = a.obs[i];
      s[0] += 1.0; s[1] += p; s[2] += p * p; s[3] += g; s[4] += p * g;
    }
  }
#pragma unroll
  for (int k = 0; k < kEvalAlignSums; ++k) {
    const double v = warp_sum_d(s[k]);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(a.scratch + k, v);
  }
}

__device__ __forceinline__ void eval_scale_shift(const double* s, float& scale, float& shift) {
  const double n = s[0], sp = s[1], spp = s[2], sg = s[3], spg = s[4];
  const double det = n * spp - sp * sp;
  const double sc = (n * spg - sp * sg) / det;
  scale = static_cast<float>(sc);
  shift = static_cast<float>((sg - sc * sp) / n);
}

__global__ void __launch_bounds__(256) eval_metric_sums_kernel(const EvalArgs a) {
  float scale, shift;
  eval_scale_shift(a.scratch, scale, shift);
  double s[2][kEvalMetricSums];
#pragma unroll
  for (int v = 0; v < 2; ++v)
#pragma unroll
    for (int k = 0; k < kEvalMetricSums; ++k) s[v][k] = 0.0;
  double cnt = 0.0;
  const int n = a.H * a.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!a.object[i]) continue;
    cnt += 1.0;
    const float p = eval_pred_at(a, i);
    const float t = __fadd_rn(a.gt[i], 1e-5f);                                   // depth_raw_ts + 1e-5 (:587)
    const float o2[2] = {__fadd_rn(p, 1e-5f), __fadd_rn(__fadd_rn(__fmul_rn(p, scale), shift), 1e-5f)};  // pred / aligned
    const float lt = logf(t), l10t = log10f(t), it = 1.0f / t;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const float o = o2[v];
      const float d = o - t, ad = fabsf(d);
      const float ld = logf(o) - lt;
      const float r = fmaxf(o / t, t / o);
      const float id = 1.0f / o - it;
      s[v][0] += ad / t;                     // abs_relative_difference  metric.py:37-47
      s[v][1] += (ad * ad) / t;              // squared_relative_difference :50-62
      s[v][2] += static_cast<double>(d) * d;        // rmse_linear :65-77
      s[v][3] += static_cast<double>(ld) * ld;      // rmse_log :80-90, silog first term :149-158
      s[v][4] += ld;                         // silog second term
      s[v][5] += fabsf(log10f(o) - l10t);    // log10 :93-101
      s[v][6] += (r < 1.25f) ? 1.0 : 0.0;    // delta1_acc :105-121 (1.25, 1.25**2, 1.25**3 as Python floats)
      s[v][7] += (r < 1.5625f) ? 1.0 : 0.0;
      s[v][8] += (r < 1.953125f) ? 1.0 : 0.0;
      s[v][9] += static_cast<double>(id) * id;      // i_rmse :132-145
    }
  }
  cnt = warp_sum_d(cnt);
  if ((threadIdx.x & 31) == 0 && cnt != 0.0) atomicAdd(a.scratch + kEvalAlignSums, cnt);
#pragma unroll
  for (int v = 0; v < 2; ++v)
#pragma unroll
    for (int k = 0; k < kEvalMetricSums; ++k) {
      const double x = warp_sum_d(s[v][k]);
      if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(a.scratch + kEvalAlignSums + 1 + v * kEvalMetricSums + k, x);
    }
}

__global__ void eval_finalize_kernel(const EvalArgs a) {
  float scale, shift;
  eval_scale_shift(a.scratch, scale, shift);
  a.out[0] = scale;
  a.out[1] = shift;
  const double n = a.scratch[kEvalAlignSums];
  for (int v = 0; v < 2; ++v) {
    const double* s = a.scratch + kEvalAlignSums + 1 + v * kEvalMetricSums;
    double* o = a.out + 2 + v * kEvalMetricSums;
    o[0] = s[0] / n;
    o[1] = s[1] / n;
    o[2] = sqrt(s[2] / n);
    o[3] = sqrt(s[3] / n);
    o[4] = s[5] / n;
    o[5] = s[6] / n;
    o[6] = s[7] / n;
    o[7] = s[8] / n;
    o[8] = sqrt(s[9] / n);
    o[9] = sqrt(s[3] / n - (s[4] * s[4]) / (n * n)) * 100.0;  // silog_rmse
  }
  a.out[22] = a.scratch[0];
  a.out[23] = n;
}

}  // namespace ada
