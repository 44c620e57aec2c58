// Bandwidth-bound kernels of the path: LayerNorm (token and channel), patch gather, cls rows,
// bilinear (align_corners=True) upsampling. All vectorised to 16-byte accesses, fp32 statistics.
#pragma once
#include "ptx.cuh"

namespace ada {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Token LayerNorm: fp32 rows [rows, D] -> bf16 rows. One warp per row, the row lives in registers (two-pass stats).
// Reference: nn.LayerNorm(D, eps=1e-6) block.py:84,87 and the shared final norm dinov2.py:337-338.
// drop_cls != 0: input rows are [B, n_tok, D]; token 0 (cls) is skipped and the output is the dense patch map
// [B, n_tok-1, D] == NHWC [B, h, w, D] (dinov2.py:339-340 + dpt.py:168-171 collapse into the store address).
// delta / delta2 != nullptr: the pending residual-branch outputs (bf16, already LayerScale'd by the GEMM epilogue) are
// added first, x <- (x + delta) + delta2 (block.py:105-106), and written back when write_x is set; the fp32 stream is
// therefore only ever touched by this coalesced kernel, never by the (row-per-thread) GEMM epilogue. The encoder writes
// x once per block: norm2 normalises x + attn-branch without storing it, the next norm1 adds both branches and stores
// (22 instead of 24 bytes per element and block).
// out2 != nullptr: a second affine of the SAME normalised row, written as the cls-less patch map -- the shared final norm of
// a tapped block (dinov2.py:337-340) has the same input, hence the same mean / variance, as the next block's norm1, so the
// tap costs one extra bf16 store instead of another pass over the fp32 stream and both pending branches.
template <int CHUNKS>  // D = CHUNKS * 128
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                      const __nv_bfloat16* __restrict__ delta2, const float* __restrict__ w, const float* __restrict__ b,
                      __nv_bfloat16* __restrict__ out, int rows, float eps, int n_tok, int drop_cls, int write_x,
                      const float* __restrict__ w2, const float* __restrict__ b2, __nv_bfloat16* __restrict__ out2) {
  constexpr int D = CHUNKS * 128;
  griddep_wait();
  griddep_launch_dependents();
  // Rows are walked from the LAST to the first: the producer of the pending branch (proj / fc2 GEMM, ascending M tiles) has
  // just finished on the highest rows, so those are the part of its 90 MB output that the 126 MB L2 still holds when this
  // kernel starts -- walking up from row 0 misses all of it and evicts it before getting there -- and this kernel in turn
  // ends on row 0, where the consuming GEMM starts. Measured at batch 32 (same box, interleaved, profiles/README.md):
  // 4.55-4.64 -> 4.26-4.44 ms per step over the 49 launches (~63 MB fewer DRAM reads per launch). Marking the fp32 stream
  // evict-first (createpolicy + ld/st .L2::cache_hint) or the GEMM's A operand (TMA .L2::cache_hint) added nothing measurable.
  const int row = rows - 1 - (blockIdx.x * 8 + (threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  if (row < 0) return;
  long long orow = row;
  if (drop_cls) {
    const int bi = row / n_tok, t = row % n_tok;
    if (t == 0) return;
    orow = static_cast<long long>(bi) * (n_tok - 1) + (t - 1);
  }
  float4* xr = reinterpret_cast<float4*>(x + static_cast<long long>(row) * D);
  float4 v[CHUNKS];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) v[i] = xr[lane + 32 * i];
  if (delta != nullptr) {
    const uint2* dr = reinterpret_cast<const uint2*>(delta + static_cast<long long>(row) * D);
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) {
      const uint2 d = dr[lane + 32 * i];
      v[i].x += bf16_lo(d.x); v[i].y += bf16_hi(d.x); v[i].z += bf16_lo(d.y); v[i].w += bf16_hi(d.y);
    }
    if (delta2 != nullptr) {
      const uint2* er = reinterpret_cast<const uint2*>(delta2 + static_cast<long long>(row) * D);
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) {
        const uint2 d = er[lane + 32 * i];
        v[i].x += bf16_lo(d.x); v[i].y += bf16_hi(d.x); v[i].z += bf16_lo(d.y); v[i].w += bf16_hi(d.y);
      }
    }
    if (write_x) {
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) xr[lane + 32 * i] = v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  uint2* o = reinterpret_cast<uint2*>(out + orow * D);
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
    const float y0 = (v[i].x - mean) * rstd * ww.x + bb.x;
    const float y1 = (v[i].y - mean) * rstd * ww.y + bb.y;
    const float y2 = (v[i].z - mean) * rstd * ww.z + bb.z;
    const float y3 = (v[i].w - mean) * rstd * ww.w + bb.w;
    o[lane + 32 * i] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
  }
  if (out2 != nullptr) {
    const int bi = row / n_tok, t = row % n_tok;
    if (t == 0) return;
    uint2* o2 = reinterpret_cast<uint2*>(out2 + (static_cast<long long>(bi) * (n_tok - 1) + (t - 1)) * D);
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w2) + lane + 32 * i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b2) + lane + 32 * i);
      const float y0 = (v[i].x - mean) * rstd * ww.x + bb.x;
      const float y1 = (v[i].y - mean) * rstd * ww.y + bb.y;
      const float y2 = (v[i].z - mean) * rstd * ww.z + bb.z;
      const float y3 = (v[i].w - mean) * rstd * ww.w + bb.w;
      o2[lane + 32 * i] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Channel LayerNorm + ReLU over NHWC bf16 pixels (channels_first LayerNorm of dpt.py:56-61 followed by nn.ReLU,
// dpt.py:156-158). One warp per pixel, C % 8 == 0, C <= 8 * 32 * MAXG. In-place safe. Persistent: every warp walks
// pixels with a grid stride, PIXELS pixels in flight per iteration, and (CACHE) keeps its lanes' weight / bias in registers
// -- the one-pixel-per-warp version re-read 16 affine scalars per 8 channels and ran at 1.45 TB/s. The kernel is latency
// bound unless enough bytes are in flight per SM (Little: ~35 KB): <1, 4> for C <= 256 (one 16-byte group per lane, four
// pixels = 2 KB per warp), <2, 2> up to 512 channels, <4, 1> / <6, 1> for 1024 / 1536.
template <int MAXG, bool CACHE, int PIXELS>
__global__ void __launch_bounds__(256)
channel_ln_relu_kernel(const __nv_bfloat16* in, const float* __restrict__ w, const float* __restrict__ b,
                       __nv_bfloat16* out, long long pixels, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int groups = C >> 3;
  const long long warps = static_cast<long long>(gridDim.x) * 8;
  float wr[CACHE ? MAXG : 1][8], br[CACHE ? MAXG : 1][8];
  if constexpr (CACHE) {
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
      const int gi = lane + 32 * i;
      if (gi < groups) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + gi * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + gi * 8) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + gi * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + gi * 8) + 1);
        wr[i][0] = w0.x; wr[i][1] = w0.y; wr[i][2] = w0.z; wr[i][3] = w0.w; wr[i][4] = w1.x; wr[i][5] = w1.y; wr[i][6] = w1.z; wr[i][7] = w1.w;
        br[i][0] = b0.x; br[i][1] = b0.y; br[i][2] = b0.z; br[i][3] = b0.w; br[i][4] = b1.x; br[i][5] = b1.y; br[i][6] = b1.z; br[i][7] = b1.w;
      }
    }
  }
  constexpr int PIX = PIXELS;  // pixels in flight per warp
  for (long long p0 = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * PIX; p0 < pixels; p0 += warps * PIX) {
    float v[PIX][MAXG][8];
    uint4 raw[PIX][MAXG];
#pragma unroll
    for (int u = 0; u < PIX; ++u) {
      const bool on = p0 + u < pixels;
      const uint4* src = reinterpret_cast<const uint4*>(in + (p0 + u) * C);
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        const int gi = lane + 32 * i;
        raw[u][i] = (on && gi < groups) ? src[gi] : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < PIX; ++u) {
      if (p0 + u >= pixels) break;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        const uint4 r = raw[u][i];
        v[u][i][0] = bf16_lo(r.x); v[u][i][1] = bf16_hi(r.x); v[u][i][2] = bf16_lo(r.y); v[u][i][3] = bf16_hi(r.y);
        v[u][i][4] = bf16_lo(r.z); v[u][i][5] = bf16_hi(r.z); v[u][i][6] = bf16_lo(r.w); v[u][i][7] = bf16_hi(r.w);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[u][i][j];  // lanes past `groups` hold zeros
      }
      const float mean = warp_sum(s) / C;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        if (lane + 32 * i < groups) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = v[u][i][j] - mean;
            q += d * d;
          }
        }
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
      uint4* dst = reinterpret_cast<uint4*>(out + (p0 + u) * C);
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        const int gi = lane + 32 * i;
        if (gi < groups) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float wj = CACHE ? wr[CACHE ? i : 0][j] : __ldg(w + gi * 8 + j);
            const float bj = CACHE ? br[CACHE ? i : 0][j] : __ldg(b + gi * 8 + j);
            y[j] = fmaxf((v[u][i][j] - mean) * rstd * wj + bj, 0.0f);
          }
          dst[gi] = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                               pack_bf16x2(y[6], y[7]));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Patch gather: fp32 NCHW image planes -> bf16 A matrix [B*P, Kpad] of the 14x14/s14 patch-embed GEMM, with the
// ImageNet normalisation of dav2.py:65 applied to the RGB planes (guide planes pass through, dav2.py:73-74).
// K index = c*196 + ky*14 + kx, matching the [D, C, 14, 14] conv weight flattened (patch_embed.py:66).
// Sources: up to 4 tensors, each [B, ch_i, H, W]; channel c of the concatenation is found by prefix sums.
struct PatchSrc {
  const float* ptr[4];
  int ch[4];
  int n;
};
// One block = up to kPgPatches horizontally adjacent patches of one patch row. Phase 1 reads the C*14 image-row segments
// they cover (each one contiguous, lanes on consecutive floats: full 128-byte requests), normalises, converts and scatters
// into a shared-memory copy of the patches' matrix rows; phase 2 writes those rows (C*196 bf16 = 392*C bytes, contiguous in
// the A matrix) with 8-byte stores, lanes on consecutive addresses. The first version (one thread per 14-float run, fourteen
// 4-byte loads at a 56-byte lane stride and seven 4-byte stores into 32 different rows per instruction) ran at 1.27 TB/s.
constexpr int kPgPatches = 32;
__global__ void __launch_bounds__(256)
patch_gather_kernel(PatchSrc src, __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int Kpad, int chunks, int npc,
                    float m0, float m1, float m2, float s0, float s1, float s2) {
  extern __shared__ __align__(16) uint8_t pg_smem[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(pg_smem);  // [patch][C*196]
  const int pw = W / 14, ph = H / 14;
  const int chunk = blockIdx.x % chunks;
  const int py = (blockIdx.x / chunks) % ph;
  const int b = blockIdx.x / (chunks * ph);
  const int px0 = chunk * npc;
  const int np = min(npc, pw - px0);  // patches of this block
  if (np <= 0) return;
  const int rowlen = C * 196;         // bf16 elements of one matrix row
  const int seg = np * 14;            // floats of one image-row segment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- phase 1: (c, ky) segments round-robin over the warps
  for (int r = warp; r < C * 14; r += 8) {
    const int c = r / 14, ky = r % 14;
    int cs = c, si = 0;
    while (si < src.n - 1 && cs >= src.ch[si]) { cs -= src.ch[si]; ++si; }
    const float* p = src.ptr[si] + ((static_cast<long long>(b) * src.ch[si] + cs) * H + (py * 14 + ky)) * W + px0 * 14;
    float mean = 0.f, sd = 1.f;
    if (c == 0) { mean = m0; sd = s0; } else if (c == 1) { mean = m1; sd = s1; } else if (c == 2) { mean = m2; sd = s2; }
    __nv_bfloat16* t = tile + r * 14;
    // all loads of the segment in flight before the first use (kPgPatches * 14 / 32 = 14 per lane): with one load per
    // loop iteration a warp had 128 bytes in flight and the kernel was latency bound at the same 1.27 TB/s as before
    float v[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const int x = lane + 32 * k;
      v[k] = (x < seg) ? p[x] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const int x = lane + 32 * k;
      if (x < seg) {
        const int pl = x / 14, kx = x - pl * 14;
        t[pl * rowlen + kx] = __float2bfloat16_rn((v[k] - mean) / sd);
      }
    }
  }
  __syncthreads();
  // ---- phase 2: matrix rows, 8 bytes per lane (392*C bytes per row: a multiple of 8, as is the row pitch Kpad*2)
  const int v8 = rowlen >> 2;
  __nv_bfloat16* o = out + (static_cast<long long>(b) * ph * pw + static_cast<long long>(py) * pw + px0) * Kpad;
  for (int i = threadIdx.x; i < np * v8; i += 256) {
    const int pl = i / v8, k = i - pl * v8;
    reinterpret_cast<uint2*>(o + static_cast<long long>(pl) * Kpad)[k] = reinterpret_cast<const uint2*>(tile + pl * rowlen)[k];
  }
}

// cls rows of the token stream: x[b, 0, :] = cls_token + pos_embed[0]  (dinov2.py:245-246)
__global__ void cls_rows_kernel(const float* __restrict__ cls_pos, float* __restrict__ x, int B, int n_tok, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i % D;
  x[static_cast<long long>(b) * n_tok * D + d] = cls_pos[d];
}

// ---------------------------------------------------------------------------------------------------------------
// Bilinear upsampling, align_corners=True (blocks.py:144, dpt.py:194), NHWC bf16. One thread = 8 channels of one output
// column, walking a strip of kUpRows output rows. Index math mirrors ATen: scale = (in-1)/(out-1) in fp32, src = scale*dst,
// i0 = (int)src, i1 = i0 + (i0 < in-1). The first version (one thread per pair of output pixels, four corner loads and six
// multiply-adds per value) was instruction bound at 3.1 TB/s (5.6 instructions per byte written). Here the horizontal
// interpolation of a source row is computed once and kept in registers while the strip walks down: with the 2x maps of the
// head every source row serves two output rows, so a value costs ~3 FMAs and half an unpack instead of 6 + 4.
constexpr int kUpRows = 16;
__device__ __forceinline__ void up_hlerp(const __nv_bfloat16* row, int x0c, int x1c, float hx, float lx, float (&o)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(row + x0c));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(row + x1c));
  const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[2 * j] = hx * bf16_lo(av[j]) + lx * bf16_lo(bv[j]);
    o[2 * j + 1] = hx * bf16_hi(av[j]) + lx * bf16_hi(bv[j]);
  }
}
__global__ void __launch_bounds__(256)
upsample_bilinear_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int Hi, int Wi, int Ho,
                         int Wo, int C, int groups_shift) {
  // grid: x over (xo, 8-channel group), y = strip of output rows, z = image: no 64-bit div/mod on the hot path
  const int groups = C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int xo = (groups_shift >= 0) ? (i >> groups_shift) : (i / groups);
  const int gi = i - xo * groups;
  if (xo >= Wo) return;
  const int b = blockIdx.z;
  const float sh = (Ho > 1) ? static_cast<float>(Hi - 1) / static_cast<float>(Ho - 1) : 0.f;
  const float sw = (Wo > 1) ? static_cast<float>(Wi - 1) / static_cast<float>(Wo - 1) : 0.f;
  const float fx = sw * xo;
  const int x0 = static_cast<int>(fx);
  const int x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
  const float lx = fx - x0, hx = 1.f - lx;
  const __nv_bfloat16* base = in + static_cast<long long>(b) * Hi * Wi * C + gi * 8;
  const int x0c = x0 * C, x1c = x1 * C;
  const long long rowpitch = static_cast<long long>(Wi) * C;
  const int ybeg = static_cast<int>(blockIdx.y) * kUpRows, yend = min(ybeg + kUpRows, Ho);
  float top[8], bot[8];
  int cy0 = -1, cy1 = -1;  // source rows currently held in top / bot
  __nv_bfloat16* dst = out + ((static_cast<long long>(b) * Ho + ybeg) * Wo + xo) * C + gi * 8;
  for (int yo = ybeg; yo < yend; ++yo) {   // (all branches below are uniform over the block: they depend on yo only)
    const float fy = sh * yo;
    const int y0 = static_cast<int>(fy);
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0);
    const float ly = fy - y0, hy = 1.f - ly;
    if (y0 != cy0) {
      if (y0 == cy1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) top[j] = bot[j];
      } else {
        up_hlerp(base + y0 * rowpitch, x0c, x1c, hx, lx, top);
      }
      cy0 = y0;
      cy1 = -1;
    }
    if (y1 != cy1) {
      if (y1 == y0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bot[j] = top[j];
      } else {
        up_hlerp(base + y1 * rowpitch, x0c, x1c, hx, lx, bot);
      }
      cy1 = y1;
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(hy * top[2 * j] + ly * bot[2 * j], hy * top[2 * j + 1] + ly * bot[2 * j + 1]);
    *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    dst += static_cast<long long>(Wo) * C;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused tail of the DPT head (dpt.py:194-195): bilinear 8h -> 14h upsample (align_corners=True) of the output_conv1 map,
// output_conv2 = conv3x3(F/2 -> 32) + ReLU + conv1x1(32 -> 1) + Sigmoid.
// A 1x1 channel contraction commutes with bilinear resampling, and a 3x3 conv is a sum over taps of shifted 1x1
// contractions, so the 32 x (F/2) contraction of EACH tap is applied at LOW resolution by a tcgen05 GEMM
// (V[pix, tap*32 + co] = sum_ci W2[co, ci, tap] * y[pix, ci], stored as fp16; 3x fewer FLOPs than the conv at 14h x 14h) and this kernel
// finishes the job: out(p) = sigmoid(w3 . relu(b2 + sum_taps bilinear(V_tap)(p + d_tap)) + b3), taps that fall outside
// the image contribute zero (the conv's zero padding). The 128-channel 14h x 14h map (2.2 GB at batch 32, re-read 9x
// through L2 by the implicit-GEMM tail, which ran at 200 TFLOP/s) is never materialised.
// One CTA = 16 x 16 output pixels; the <= 12 x 12 low-resolution patch of V (288 channels) it needs is staged in shared
// memory once. Index math mirrors ATen: scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = (int)src, i1 = i0+(i0<in-1).
constexpr int kTailTile = 16;
constexpr int kTailPatch = 12;
constexpr int kTailCh = 288;  // 9 taps x 32
constexpr int kTailPitch = kTailCh * 2 + 16;  // 592 B per low-res pixel: +16 B skews neighbouring pixels across all banks
constexpr int kTailSmemBytes = kTailPatch * kTailPatch * kTailPitch;

__global__ void __launch_bounds__(256)
tail_gather_kernel(const __half* __restrict__ V, const float* __restrict__ bias2, const float* __restrict__ aux,
                   float* __restrict__ out, int Hl, int Wl, int H, int W, int apply_sigmoid) {
  extern __shared__ __align__(16) uint8_t tail_smem[];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int Y0 = blockIdx.y * kTailTile, X0 = blockIdx.x * kTailTile;
  const float sh = (H > 1) ? static_cast<float>(Hl - 1) / static_cast<float>(H - 1) : 0.f;
  const float sw = (W > 1) ? static_cast<float>(Wl - 1) / static_cast<float>(W - 1) : 0.f;
  // low-resolution patch covering every tap position of this tile
  const int r0 = static_cast<int>(sh * max(Y0 - 1, 0));
  const int r1 = min(static_cast<int>(sh * min(Y0 + kTailTile, H - 1)) + 1, Hl - 1);
  const int c0 = static_cast<int>(sw * max(X0 - 1, 0));
  const int c1 = min(static_cast<int>(sw * min(X0 + kTailTile, W - 1)) + 1, Wl - 1);
  const int nrows = r1 - r0 + 1, ncols = c1 - c0 + 1;
  if (nrows > kTailPatch || ncols > kTailPatch) {  // geometry other than 8h -> 14h: fail loudly
    if (tid == 0) g_dev_error[0] = 0x7A11;
    __trap();
  }
  {
    const int total = nrows * ncols * (kTailCh / 8);
    const uint4* src = reinterpret_cast<const uint4*>(V);
    for (int i = tid; i < total; i += 256) {
      const int pix = i / (kTailCh / 8), q = i - pix * (kTailCh / 8);
      const int r = pix / ncols, c = pix - r * ncols;
      *reinterpret_cast<uint4*>(tail_smem + pix * kTailPitch + q * 16) =
          __ldg(src + ((static_cast<long long>(b) * Hl + r0 + r) * Wl + c0 + c) * (kTailCh / 8) + q);
    }
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  const int Y = Y0 + ty, X = X0 + tx;
  if (Y >= H || X >= W) return;
  // The 9 x 4 x 32 interpolation FMAs per pixel are what bounds this kernel (issue slots, not HBM). The tap map is stored as
  // fp16, so the four corners of the three taps of one kernel row are combined by packed fp16 FMAs (HFMA2: two channels
  // per instruction, no unpacking; 12 terms of magnitude O(1) per fp16 sum), and the three row sums are added in fp32.
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = __ldg(bias2 + i);
  // per-tap source columns (three of them), rows handled in the loop
  int xo0[3], xo1[3];
  float lxv[3];
  bool xok[3];
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    const int px = X + kx - 1;
    xok[kx] = (px >= 0) && (px < W);
    const float fx = sw * (xok[kx] ? px : 0);
    const int x0 = static_cast<int>(fx);
    xo0[kx] = x0 - c0;
    xo1[kx] = x0 + (x0 < Wl - 1 ? 1 : 0) - c0;
    lxv[kx] = fx - x0;
  }
#pragma unroll 1
  for (int ky = 0; ky < 3; ++ky) {
    const int py = Y + ky - 1;
    if (py < 0 || py >= H) continue;
    const float fy = sh * py;
    const int y0 = static_cast<int>(fy);
    const int yr0 = y0 - r0, yr1 = y0 + (y0 < Hl - 1 ? 1 : 0) - r0;
    const float ly = fy - y0, hy = 1.f - ly;
    __half2 racc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) racc[i] = __float2half2_rn(0.f);
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      if (!xok[kx]) continue;
      const float lx = lxv[kx], hx = 1.f - lx;
      const int tap = ky * 3 + kx;
      const float wgt[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
      const int pidx[4] = {yr0 * ncols + xo0[kx], yr0 * ncols + xo1[kx], yr1 * ncols + xo0[kx], yr1 * ncols + xo1[kx]};
#pragma unroll
      for (int cnr = 0; cnr < 4; ++cnr) {
        const uint4* p = reinterpret_cast<const uint4*>(tail_smem + pidx[cnr] * kTailPitch + tap * 64);
        const __half2 w2 = __float2half2_rn(wgt[cnr]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 v = p[q];
          racc[q * 4 + 0] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v.x), racc[q * 4 + 0]);
          racc[q * 4 + 1] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v.y), racc[q * 4 + 1]);
          racc[q * 4 + 2] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v.z), racc[q * 4 + 2]);
          racc[q * 4 + 3] = __hfma2(w2, *reinterpret_cast<const __half2*>(&v.w), racc[q * 4 + 3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 f = __half22float2(racc[i]);
      acc[2 * i] += f.x;
      acc[2 * i + 1] += f.y;
    }
  }
  float sres = __ldg(aux + 32);
#pragma unroll
  for (int i = 0; i < 32; ++i) sres = fmaf(fmaxf(acc[i], 0.f), __ldg(aux + i), sres);
  if (apply_sigmoid == 1) sres = 1.0f / (1.0f + __expf(-sres));
  else if (apply_sigmoid == 2) sres = fmaxf(sres, 0.f);  // un-guided head ends in ReLU (depth_anything_v2_raw/dpt.py:115,182)
  out[(static_cast<long long>(b) * H + Y) * W + X] = sres;
}

}  // namespace ada
