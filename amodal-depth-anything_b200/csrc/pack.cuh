// Weight packing on the device: fp32 state-dict tensors (staged in device memory by ada_set_weight) -> the bf16 operand
// layouts of the kernels. One gather-and-cast kernel, the source index computed from the destination index per layout.
// Replaces the round-1 path (every tensor copied device -> host, repacked by scalar host loops, copied back: ~8 GB of PCIe
// traffic and several seconds for ViT-G, repeated on every .to() / load_state_dict of the Python module).
#pragma once
#include "ptx.cuh"

namespace ada {

enum PackMode : int {
  PACK_CAST = 0,     // dst[i] = src[i]                                                  (Linear / 1x1 conv weights)
  PACK_CONV3X3 = 1,  // [Cout,Cin,3,3] -> [Cout, 9*Cpad], k = tap*Cpad + ci, zero padded  (dpt.py / blocks.py 3x3 convs)
  PACK_CONVT = 2,    // ConvTranspose2d [Cin,Cout,ks,ks], stride == ks -> [(kk*Cout + co), Cin]      (dpt.py:89-100)
  PACK_TAIL = 3,     // output_conv2.0 [32,Cm,3,3] -> per-tap 1x1 contractions [(tap*32 + co), Cm]    (dpt.py:146-151)
  PACK_EMBED = 4,    // patch_embed [D,3,14,14] ++ patch_embed_guidance [D,Cg,14,14] -> [D, Kpad]     (dinov2.py:234-240)
  PACK_SWIGLU = 5    // w12 [2*Hd, D]: rows interleaved in 32-row (x1, x2) chunk pairs               (swiglu_ffn.py:30-32)
};
struct PackDesc {
  int mode;
  int a, b, c, d;  // CONV3X3: Cout, Cin, Cpad | CONVT: Cin, Cout, ks | TAIL: Cm | EMBED: D, Cg, Kpad | SWIGLU: Hd, D
};

__device__ __forceinline__ long long swiglu_src_row(long long r, int Hd) {
  const long long chunk = r / 64, within = r % 64;
  return (within < 32) ? chunk * 32 + within : Hd + chunk * 32 + (within - 32);
}

__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ src, const float* __restrict__ src2, __nv_bfloat16* __restrict__ dst,
                    long long n, PackDesc p) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  switch (p.mode) {
    case PACK_CAST: v = src[i]; break;
    case PACK_CONV3X3: {
      const int Cin = p.b, Cpad = p.c;
      const long long co = i / (9LL * Cpad);
      const int rem = static_cast<int>(i - co * 9LL * Cpad);
      const int t = rem / Cpad, ci = rem - t * Cpad;
      if (ci < Cin) v = src[(co * Cin + ci) * 9 + t];
      break;
    }
    case PACK_CONVT: {
      const int Cin = p.a, Cout = p.b, ks2 = p.c * p.c;
      const long long kk = i / (static_cast<long long>(Cout) * Cin);
      const long long rem = i - kk * Cout * Cin;
      const long long co = rem / Cin, ci = rem - co * Cin;
      v = src[(ci * Cout + co) * ks2 + kk];
      break;
    }
    case PACK_TAIL: {
      const int Cm = p.a;
      const long long t = i / (32LL * Cm);
      const long long rem = i - t * 32LL * Cm;
      const long long co = rem / Cm, ci = rem - co * Cm;
      v = src[(co * Cm + ci) * 9 + t];
      break;
    }
    case PACK_EMBED: {
      const int Cg = p.b, Kpad = p.c;
      const long long d = i / Kpad;
      const int k = static_cast<int>(i - d * Kpad);
      if (k < 588) v = src[d * 588 + k];
      else if (k < 588 + Cg * 196) v = src2[d * Cg * 196 + (k - 588)];
      break;
    }
    case PACK_SWIGLU: {
      const int Hd = p.a, D = p.b;
      const long long r = i / D, k = i - r * D;
      v = src[swiglu_src_row(r, Hd) * D + k];
      break;
    }
  }
  dst[i] = __float2bfloat16_rn(v);
}

// fp32 row permutation of the SwiGLU bias (same interleave as the weight rows)
__global__ void swiglu_bias_kernel(const float* __restrict__ src, float* __restrict__ dst, int Hd) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < 2 * Hd) dst[r] = src[swiglu_src_row(r, Hd)];
}

}  // namespace ada
