// Tail of the DPT head on tensor cores (dpt.py:194-195): bilinear 8h -> 14h upsample (align_corners=True) of the
// output_conv1 map, output_conv2 = conv3x3(C -> 32) + ReLU + conv1x1(32 -> 1) + Sigmoid, in ONE kernel that reads the
// C-channel low-resolution map once (fp16, written by output_conv1's EPI_F16 epilogue) and writes 4 bytes per pixel.
// Neither the upsampled [B, 14h, 14w, C] map (2.2 GB at batch 32) nor a per-tap map ([B, 8h, 8w, 288]: 1.6 GB written by a
// GEMM at 5 TB/s and re-read by tail_gather_kernel at 1.2 TB/s, 1.8 ms together) exists.
//
// Persistent CTAs, one tile = 8 x 14 output pixels:
//   warp 1      TMA: the 8 x 12 low-resolution patch the tile needs (all C channels) into a 2-deep ring
//   warps 8-15  interpolation: thread = (column of the 10 x 16 halo tile, 8-channel chunk); horizontal then vertical lerp as
//               packed fp16 FMAs (a + t (b - a)), pixels outside the image are zero (the conv's padding); the tile is written
//               in the NO-SWIZZLE K-major core-matrix layout (8 pixels x 16 bytes contiguous, chunk planes LBO apart)
//   warp 0      tcgen05: D[q, kx*32 + co] (+)= U[q + 16 ky, ci] W[ky][ci, kx*32 + co], M = 128 halo-tile pixels (8 rows x 16
//               columns, linear, so a kernel-row shift is a 256-byte start-address shift), N = 96 = the three taps of a
//               kernel row side by side, K = 16 channels per instruction: 3 x C/16 instructions per tile. With the kx taps in
//               N the A tile is read 3 times instead of 9 (a 128 x 32 x 16 MMA is bound by its 4 KB A read, 40 cycles for 16
//               cycles of math).
//   warps 4-7   epilogue: thread = pixel q; out(y, x) = D[q, 0:32] + D[q+1, 32:64] + D[q+2, 64:96] (two warp shuffles per
//               channel; columns 14, 15 of the halo tile produce nothing), + bias, ReLU, 32-wide dot, sigmoid, fp32 store.
// Index math mirrors ATen: scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = (int)src, i1 = i0+(i0<in-1).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace ada {

constexpr int kTmTileH = 8;                      // output rows per tile
constexpr int kTmTileW = 14;                     // output columns per tile (W is a multiple of 14: whole patches)
constexpr int kTmUW = 16;                        // halo tile: columns X0-1 .. X0+14
constexpr int kTmUH = 10;                        // halo tile: rows Y0-1 .. Y0+8
constexpr int kTmULbo = kTmUW * kTmUH * 16 + 16; // bytes between 8-channel chunk planes (+16: bank skew for the writers)
constexpr int kTmPatchH = 8, kTmPatchW = 12;     // low-resolution patch (9 / 15 output steps of < 4/7 plus the +1 neighbour)
constexpr int kTmThreads = 512;
constexpr int kTmTmemCols = 256;                 // two accumulators of 96 columns at 0 and 128
constexpr int kTmMaxC = 128;

struct TailMmaArgs {
  const __half* wpk;   // [3 ky][C/8 chunks][96 = kx*32 + co][8 ci] fp16 (pack_tail_mma_kernel)
  const float* bias2;  // [32] output_conv2.0 bias
  const float* aux;    // [33] output_conv2.2 weight + bias
  float* out;          // [B, H, W] fp32
  int B, Hl, Wl, H, W, C, sigmoid;
  int tiles_x, tiles_y, total_tiles;
  uint32_t magic_x, magic_y;  // ceil(2^32 / tiles_x), ceil(2^32 / tiles_y): exact division by multiply-high for t * d < 2^32
};

__host__ __device__ constexpr int tm_round128(int x) { return (x + 127) & ~127; }
__host__ __device__ constexpr int tm_u_bytes(int C) { return (C / 8) * kTmULbo; }
__host__ __device__ constexpr int tm_l_bytes(int C) { return kTmPatchH * kTmPatchW * C * 2; }
__host__ __device__ constexpr int tm_off_u(int C) { return 576 * C; }
__host__ __device__ constexpr int tm_off_l(int C) { return tm_round128(tm_off_u(C) + 2 * tm_u_bytes(C)); }
__host__ __device__ constexpr int tm_off_bar(int C) { return tm_off_l(C) + 2 * tm_l_bytes(C); }
__host__ __device__ constexpr int tm_off_aux(int C) { return tm_off_bar(C) + 128; }
__host__ __device__ constexpr int tm_off_info(int C) { return tm_off_aux(C) + 384; }  // 8 tile-info slots of 128 bytes
__host__ __device__ constexpr int tm_smem_bytes(int C) { return tm_off_info(C) + 8 * 128; }

// Shared-memory matrix descriptor without swizzle (K-major): core matrix = 8 rows x 16 bytes, contiguous;
// lbo = bytes between the two 16-byte K chunks of one instruction, sbo = bytes between 8-row groups.
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  return d;
}
// Instruction descriptor: D = f32, A = B = fp16, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// a + t * (b - a) on 8 packed fp16 values
__device__ __forceinline__ uint4 lerp_h8(const uint4& a, const uint4& b, __half2 t) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hfma2(t, __hsub2(pb[i], pa[i]), pa[i]);
  return r;
}

// Vertical pass for halo row YY of one thread's column: h[] = the column's 8 horizontally interpolated patch rows.
// Relative to the tile's first source row, halo row YY reads source row d = floor(YY * s), s = (8h - 1) / (14h - 1) just under
// 4/7, or the one after it -- which of the two depends on the fractional position of the tile and is the same for every
// thread, so the TMA warp publishes it as one bit per row and the register pair is picked by selects (no branches, no
// dynamic register index). Tiles with a halo row outside the image (first / last tile row) take the general path below.
__host__ __device__ constexpr int tm_row_d(int yy) { return (yy * 57) / 100; }
template <int YY>
__device__ __forceinline__ void tm_vrow_fast(const uint4 (&h)[kTmPatchH], uint32_t mask, uint32_t lyb, uint32_t up) {
  constexpr int d = tm_row_d(YY);
  static_assert(d + 2 < kTmPatchH, "row table exceeds the patch");
  const bool p = (mask >> YY) & 1u;
  const uint4 lo = p ? h[d + 1] : h[d], hi = p ? h[d + 2] : h[d + 1];
  const uint4 v = lerp_h8(lo, hi, *reinterpret_cast<const __half2*>(&lyb));
  st_shared_v4(up + static_cast<uint32_t>(YY) * (kTmUW * 16), v.x, v.y, v.z, v.w);
}
template <int YY>
__device__ __forceinline__ void tm_vrow(const uint4 (&h)[kTmPatchH], uint32_t slot, uint32_t up) {
  int jrow;
  uint32_t lyb;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(jrow) : "r"(slot + 32u + 4u * YY));
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lyb) : "r"(slot + 80u + 4u * YY));
  const __half2 ly2 = *reinterpret_cast<const __half2*>(&lyb);
  uint4 v;
  switch (jrow) {
    case 0: v = lerp_h8(h[0], h[1], ly2); break;
    case 1: v = lerp_h8(h[1], h[2], ly2); break;
    case 2: v = lerp_h8(h[2], h[3], ly2); break;
    case 3: v = lerp_h8(h[3], h[4], ly2); break;
    case 4: v = lerp_h8(h[4], h[5], ly2); break;
    case 5: v = lerp_h8(h[5], h[6], ly2); break;
    case 6: v = lerp_h8(h[6], h[7], ly2); break;
    case 7: v = h[7]; break;  // the last image row only (checked by the TMA warp): the row below has weight zero
    default: v = make_uint4(0, 0, 0, 0); break;  // halo row outside the image
  }
  st_shared_v4(up + static_cast<uint32_t>(YY) * (kTmUW * 16), v.x, v.y, v.z, v.w);
}

__global__ void __launch_bounds__(kTmThreads, 1)
tail_mma_kernel(const __grid_constant__ CUtensorMap tmap_l, const TailMmaArgs a) {
  extern __shared__ __align__(1024) uint8_t tm_smem[];
  const uint32_t sbase = smem_u32(tm_smem);
  const int C = a.C, nch = C >> 3;
  const uint32_t u_bytes = static_cast<uint32_t>(tm_u_bytes(C)), l_bytes = static_cast<uint32_t>(tm_l_bytes(C));
  const uint32_t sW = sbase, sU = sbase + tm_off_u(C), sL = sbase + tm_off_l(C), bar = sbase + tm_off_bar(C);
  float* s_aux = reinterpret_cast<float*>(tm_smem + tm_off_aux(C));  // [0:32] bias2, [32:64] w3, [64] b3
  auto l_full = [&](int b) { return bar + 8u * b; };
  auto l_empty = [&](int b) { return bar + 16u + 8u * b; };
  auto u_full = [&](int b) { return bar + 32u + 8u * b; };
  auto u_empty = [&](int b) { return bar + 48u + 8u * b; };
  auto acc_full = [&](int b) { return bar + 64u + 8u * b; };
  auto acc_empty = [&](int b) { return bar + 80u + 8u * b; };
  const uint32_t tmem_ptr_smem = bar + 96u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (sbase & 127u) {
      g_dev_error[0] = 0x7A20;
      __trap();
    }
    tma_prefetch_desc(&tmap_l);
    for (int b = 0; b < 2; ++b) {
      mbar_init(l_full(b), 1);
      mbar_init(l_empty(b), 8);    // one arrival per interpolation warp
      mbar_init(u_full(b), 8);
      mbar_init(u_empty(b), 1);    // tcgen05.commit
      mbar_init(acc_full(b), 1);   // tcgen05.commit
      mbar_init(acc_empty(b), 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kTmTmemCols);
    tmem_relinquish();
  }
  {  // weights (constants: not produced by the previous kernel, so this runs ahead of the dependency wait)
    const uint4* src = reinterpret_cast<const uint4*>(a.wpk);
    const int n16 = 36 * C;  // 576 C bytes
    for (int i = threadIdx.x; i < n16; i += kTmThreads) {
      const uint4 v = __ldg(src + i);
      st_shared_v4(sW + 16u * i, v.x, v.y, v.z, v.w);
    }
    if (threadIdx.x < 32) s_aux[threadIdx.x] = __ldg(a.bias2 + threadIdx.x);
    else if (threadIdx.x < 65) s_aux[threadIdx.x] = __ldg(a.aux + threadIdx.x - 32);
    else if (threadIdx.x < 72) s_aux[threadIdx.x + 3] = 0.f;  // floats 68..74: the 16 zero bytes at byte offset 272
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  const float sh = (a.H > 1) ? static_cast<float>(a.Hl - 1) / static_cast<float>(a.H - 1) : 0.f;
  const float sw = (a.W > 1) ? static_cast<float>(a.Wl - 1) / static_cast<float>(a.W - 1) : 0.f;
  // Tile parameters are computed once per tile by the TMA warp and published in an 8-deep ring of 128-byte slots
  // (words [0] img, [1] Y0, [2] X0, [3] r0, [4] c0, [5] fast-path flag, [6] row-select bits; [8 + yy] source-row index of
  // halo row yy relative to r0 (-1: outside the image); [20 + yy] its vertical weight as a packed fp16 pair), so the 384 worker threads do not repeat the divisions
  // and float <-> int conversions. A slot is rewritten 8 tiles later, by which time every reader of it has moved on (the
  // load of tile k+8 needs the interpolation of tile k+6, that the MMAs of tile k+3, those the epilogue of tile k+1).
  const uint32_t s_info = sbase + tm_off_info(C);

  if (warp == 1) {
    // ------------------------------------------------------------------ TMA: low-resolution patches + tile parameters
    int k = 0;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++k) {
      const int b = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      // (2^32 / 1 does not fit the 32-bit magic number: a divisor of one is passed through)
      const uint32_t r = a.tiles_x == 1 ? static_cast<uint32_t>(t) : __umulhi(static_cast<uint32_t>(t), a.magic_x);  // t / tiles_x
      const int tx = t - static_cast<int>(r) * a.tiles_x;
      const int img = static_cast<int>(a.tiles_y == 1 ? r : __umulhi(r, a.magic_y));  // r / tiles_y
      const int ty = static_cast<int>(r) - img * a.tiles_y;
      const int Y0 = ty * kTmTileH, X0 = tx * kTmTileW;
      const int r0 = static_cast<int>(sh * static_cast<float>(max(Y0 - 1, 0)));
      const int c0 = static_cast<int>(sw * static_cast<float>(max(X0 - 1, 0)));
      const uint32_t slot = s_info + 128u * (k & 7);
      int jrow = -1;
      uint32_t lyb = 0;
      if (lane < kTmUH) {
        const int Y = Y0 - 1 + lane;
        if (Y >= 0 && Y < a.H) {
          const float fy = sh * static_cast<float>(Y);
          const int y0 = static_cast<int>(fy);
          jrow = y0 - r0;
          const __half2 ly2 = __float2half2_rn(y0 < a.Hl - 1 ? fy - static_cast<float>(y0) : 0.f);
          lyb = *reinterpret_cast<const uint32_t*>(&ly2);
          if (jrow < 0 || jrow > kTmPatchH - 1 || (jrow == kTmPatchH - 1 && y0 < a.Hl - 1)) {  // geometry other than 8h -> 14h
            g_dev_error[0] = 0x7A22;
            __trap();
          }
        }
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(slot + 32u + 4u * lane), "r"(jrow) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(slot + 80u + 4u * lane), "r"(lyb) : "memory");
      }
      const int drow = tm_row_d(lane);
      const uint32_t in_table = __ballot_sync(0xffffffffu, lane >= kTmUH || jrow == drow || jrow == drow + 1);
      const uint32_t sel_bits = __ballot_sync(0xffffffffu, lane < kTmUH && jrow == drow + 1);
      if (lane == 0) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(img), "r"(Y0), "r"(X0), "r"(r0) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 16u), "r"(c0), "r"(in_table == 0xffffffffu ? 1 : 0),
                     "r"(sel_bits), "r"(0) : "memory");
      }
      __syncwarp();
      mbar_wait(l_empty(b), ph ^ 1u, 0x7A0 + b);
      mbar_expect_tx_w(l_full(b), l_bytes);  // (release) publishes the slot to whoever sees this phase complete
      tma_load_4d_w(sL + b * l_bytes, &tmap_l, l_full(b), 0, c0, r0, img);
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ tcgen05 issuer
    constexpr uint32_t idesc = make_idesc_f16(128, 96);
    const int ksteps = C >> 4;
    int k = 0;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++k) {
      const int b = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      mbar_wait(u_full(b), ph, 0x7A2 + b);
      mbar_wait(acc_empty(b), ph ^ 1u, 0x7A4 + b);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 128u * b;
      const uint32_t ub = sU + b * u_bytes;
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
        for (int s = 0; s < ksteps; ++s) {
          const uint64_t ad = make_smem_desc_nosw(ub + static_cast<uint32_t>(2 * s) * kTmULbo + static_cast<uint32_t>(ky) * (kTmUW * 16), kTmULbo, 128);
          const uint64_t bd = make_smem_desc_nosw(sW + static_cast<uint32_t>(ky * nch + 2 * s) * 1536u, 1536, 128);
          umma_bf16_ss_w(d_tmem, ad, bd, idesc, (ky | s) ? 1u : 0u);
        }
      }
      umma_commit_w(u_empty(b));
      umma_commit_w(acc_full(b));
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ interpolation: halo tile into the A-operand layout
    const int pt = threadIdx.x - 256;
    const int chunk = pt % nch, xq = pt / nch;
    const bool active = xq < kTmUW;
    const uint32_t pix_b = static_cast<uint32_t>(C) * 2u, row_b = pix_b * kTmPatchW;
    const uint32_t zero16 = sbase + tm_off_aux(C) + 272u;  // 16 zero bytes (columns outside the image read these)
    int k = 0;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++k) {
      const int b = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      const uint32_t slot = s_info + 128u * (k & 7);
      mbar_wait(l_full(b), ph, 0x7A6 + b);
      // ---- horizontal pass: this thread's column of all 8 patch rows, in registers (16 independent loads in flight)
      uint4 h[kTmPatchH];
      int c0, fast;
      uint32_t sel_bits, pad_;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(c0), "=r"(fast), "=r"(sel_bits), "=r"(pad_) : "r"(slot + 16u));
      {
        int X0;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(X0) : "r"(slot + 8u));
        const int X = X0 - 1 + xq;
        const bool xok = active && X >= 0 && X < a.W;
        const float fx = sw * static_cast<float>(min(max(X, 0), a.W - 1));
        const int x0 = static_cast<int>(fx);
        const int xo0 = x0 - c0, xo1 = x0 + (x0 < a.Wl - 1 ? 1 : 0) - c0;
        const __half2 lx2 = __float2half2_rn(fx - static_cast<float>(x0));
        if (xok && (xo0 < 0 || xo1 >= kTmPatchW)) {  // geometry other than 8h -> 14h: fail loudly
          g_dev_error[0] = 0x7A21;
          __trap();
        }
        const uint32_t lp = sL + b * l_bytes + static_cast<uint32_t>(chunk) * 16u;
        // outside the image (the conv's zero padding): both loads of every row read the 16 zero bytes
        const uint32_t p0 = xok ? lp + static_cast<uint32_t>(xo0) * pix_b : zero16;
        const uint32_t p1 = xok ? lp + static_cast<uint32_t>(xo1) * pix_b : zero16;
        const uint32_t rstep = xok ? row_b : 0u;
#pragma unroll
        for (int r = 0; r < kTmPatchH; ++r) {
          const uint4 va = ld_shared_v4(p0 + static_cast<uint32_t>(r) * rstep);
          const uint4 vb = ld_shared_v4(p1 + static_cast<uint32_t>(r) * rstep);
          h[r] = lerp_h8(va, vb, lx2);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(l_empty(b));  // the patch is in registers: the next TMA may land
      mbar_wait(u_empty(b), ph ^ 1u, 0x7A8 + b);
      // ---- vertical pass. The source row of an output row is the same for every thread of the CTA, so picking the
      // register pair is a warp-uniform switch, not a dynamic register index.
      if (active) {
        const uint32_t up = sU + b * u_bytes + static_cast<uint32_t>(chunk) * kTmULbo + static_cast<uint32_t>(xq) * 16u;
        if (fast) {
          const uint4 w0 = ld_shared_v4(slot + 80u), w1 = ld_shared_v4(slot + 96u);
          uint32_t w8, w9;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w8), "=r"(w9) : "r"(slot + 112u));
          tm_vrow_fast<0>(h, sel_bits, w0.x, up); tm_vrow_fast<1>(h, sel_bits, w0.y, up);
          tm_vrow_fast<2>(h, sel_bits, w0.z, up); tm_vrow_fast<3>(h, sel_bits, w0.w, up);
          tm_vrow_fast<4>(h, sel_bits, w1.x, up); tm_vrow_fast<5>(h, sel_bits, w1.y, up);
          tm_vrow_fast<6>(h, sel_bits, w1.z, up); tm_vrow_fast<7>(h, sel_bits, w1.w, up);
          tm_vrow_fast<8>(h, sel_bits, w8, up); tm_vrow_fast<9>(h, sel_bits, w9, up);
        } else {
          tm_vrow<0>(h, slot, up); tm_vrow<1>(h, slot, up); tm_vrow<2>(h, slot, up); tm_vrow<3>(h, slot, up);
          tm_vrow<4>(h, slot, up); tm_vrow<5>(h, slot, up); tm_vrow<6>(h, slot, up); tm_vrow<7>(h, slot, up);
          tm_vrow<8>(h, slot, up); tm_vrow<9>(h, slot, up);
        }
        static_assert(kTmUH == 10, "ten halo rows");
      }
      fence_proxy_async_smem();  // the tile is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(u_full(b));
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int e = warp - 4;
    const int q = e * 32 + lane, y = q >> 4, x = q & 15;
    const float b3 = s_aux[64];
    int k = 0;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++k) {
      const int b = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      mbar_wait(acc_full(b), ph, 0x7AA + b);
      tc_fence_after();
      const uint32_t taddr = tmem_base + 128u * b + (static_cast<uint32_t>(e * 32) << 16);
      uint32_t d0[32], d1[32];
      tmem_ld32(taddr, d0);
      tmem_ld32(taddr + 32, d1);
      int img, Y0, X0;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(img) : "r"(s_info + 128u * (k & 7)));
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(Y0) : "r"(s_info + 128u * (k & 7) + 4u));
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(X0) : "r"(s_info + 128u * (k & 7) + 8u));
      tmem_ld_wait();
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(d0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(d1[i]), 1);
      tmem_ld32(taddr + 64, d0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(b));
      float sres = b3;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(s_aux + i), ww = *reinterpret_cast<const float4*>(s_aux + 32 + i);
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = acc[i + j] + __shfl_down_sync(0xffffffffu, __uint_as_float(d0[i + j]), 2) + bv[j];
          sres = fmaf(fmaxf(v, 0.f), wv[j], sres);
        }
      }
      if (a.sigmoid == 1) sres = 1.0f / (1.0f + __expf(-sres));
      else if (a.sigmoid == 2) sres = fmaxf(sres, 0.f);  // un-guided head ends in ReLU (depth_anything_v2_raw/dpt.py:115,182)
      const int Y = Y0 + y;
      if (x < kTmTileW && Y < a.H) a.out[(static_cast<long long>(img) * a.H + Y) * a.W + X0 + x] = sres;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmTmemCols);
  }
}

// output_conv2.0 weight [32, C, 3, 3] fp32 -> [ky][C/8][kx*32 + co][8] fp16: the B operand of tail_mma_kernel, already in
// the no-swizzle core-matrix order so the kernel copies it linearly.
__global__ void pack_tail_mma_kernel(const float* __restrict__ w, __half* __restrict__ dst, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 288 * C) return;
  const int e = i & 7, n = (i >> 3) % 96, rest = (i >> 3) / 96;
  const int nch = C >> 3, c = rest % nch, ky = rest / nch;
  const int kx = n >> 5, co = n & 31, ci = c * 8 + e;
  dst[i] = __float2half_rn(w[((co * C + ci) * 3 + ky) * 3 + kx]);
}

}  // namespace ada
