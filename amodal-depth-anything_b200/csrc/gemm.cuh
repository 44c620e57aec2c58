// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] * B[N,K]^T).
//
//   warp 0      : TMA producer  (A and B tiles, 128-byte swizzle, 4..8 stage mbarrier ring)
//   warp 1      : MMA issuer    (tcgen05.mma, 128 x BN x 16 per CTA, fp32 accumulators in TMEM, 2 accumulator stages)
//                 Both run their loops with the whole warp converged and ONE elect.sync-predicated lane per TMA / tcgen05
//                 instruction (ptx.cuh *_w wrappers): inside `if (lane == 0)` the compiler builds an ELECT / R2UR /
//                 BRA.U.ANY waterfall around every descriptor (13-22 instructions per MMA) and, competing for issue slots
//                 with the epilogue warps of the same scheduler, one MMA issue took ~160 cycles -- longer than the MMA.
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue      (tcgen05.ld 32x32b -> registers -> fused epilogue -> swizzled smem -> TMA store);
//                                 two warpgroups split the tile's 64-column groups: ncu showed the 4-warp epilogue
//                                 latency-bound (issue 26%, XU 25%, tensor pipe 50% on the GELU GEMM), not pipe-bound.
//                                 setmaxnreg moves registers from the producer/MMA warpgroup to the epilogue warpgroups.
//
// CG = 2 runs the kernel as CTA pairs (cluster of 2, tcgen05 cta_group::2): the pair computes a 256 x BN tile, each CTA
// loads its own 128 rows of A but only HALF of the B tile, and the leader CTA issues one 256 x BN x 16 MMA that reads
// both halves. Operand traffic from L2 per FLOP drops by a third (the 128 x 256 single-CTA tile was L2->SM bound at
// ~1.35 PFLOP/s) and shared-memory reads of B halve. Barriers: both producers credit the leader's `full` barrier,
// tcgen05.commit multicasts `empty` / `tmem-full` to both CTAs, both epilogues release the leader's `tmem-empty`.
//
// The A operand is fetched either as a plain row-major matrix (linear layers, 1x1 convs, transposed convs with k == s)
// or as an implicit-GEMM 3x3/pad-1 convolution (stride 1 or 2) over an NHWC map: the M tile is an 8x16 patch of OUTPUT
// pixels and each of the 9 taps is one shifted 4-D TMA box whose out-of-bounds part the hardware zero-fills (that is the
// padding); for stride 2 the tensor map traverses the input with element stride 2 in x and y.
//
// Epilogue data path. After tcgen05.ld each thread owns one accumulator row; storing that straight to global touches
// 32 different cache lines per warp instruction (measured: 32 sectors/request, K=1024 GEMMs ran epilogue-bound at
// ~45% of the K=4096 rate). Instead every epilogue warp converts 64 columns at a time, writes its 32x64 bf16 block
// into a private 4 KB swizzled staging buffer (conflict-free 16-byte stores) and one lane issues a TMA store; the
// hardware clips rows/columns/pixels that fall outside the tensor, so ragged M, N, H, W need no predication.
// The arithmetic is packed fp32 (fma.rn.f32x2) and activation / residual handling are template parameters (EpiMode), so
// the plain path costs 1.6 instructions per output. Shared-memory bandwidth is the scarce resource of the 256-wide main
// loop (~125 B/clk of TMA writes + MMA reads): epilogue staging traffic is kept to 4 KB per 64 columns and warp.
//
// Reference ops this kernel replaces (all fp32 torch ops in the reference):
//   attention.py:51,60  mlp.py:36-39  swiglu_ffn.py:30-33  layer_scale.py:28                     (encoder linears)
//   patch_embed.py:76   dinov2.py:234-246                                                       (patch embed + pos)
//   dpt.py:172-173,178  blocks.py:20-24,57-80,146  dpt.py:193,195                                (DPT head convs)
#pragma once
#include "ptx.cuh"

namespace ada {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kTileH = 8;     // conv mode: M tile = 8 x 16 pixels
constexpr int kTileW = 16;
constexpr int kGemmThreads = 384;
constexpr int kEpiThreads = 256;   // warps 4..11

enum EpiMode : int {
  EPI_BF16 = 0,       // out_bf16 = act((acc + bias) * gamma) [+ resid1 + resid2]; optional second copy with ReLU; TMA store
  EPI_EMBED = 2,      // out_f32[b*(P+1)+1+p, :] = acc + aux[p, :]           (patch embed: + bias + pos-embed, skip cls row)
  EPI_CONVT = 3,      // k == s transposed conv: pixel-shuffle scatter, + bias[co]
  EPI_TAIL = 4,       // sigmoid(relu(acc + bias) . aux[0:32] + aux[32]) -> fp32 per pixel (BN must be 32)
  EPI_SWIGLU = 5,     // columns interleaved in 32-wide (x1, x2) chunk pairs: out = silu(x1 + b1) * (x2 + b2); TMA store
  // compile-time specialisations of EPI_BF16, picked by the launcher from (act, residuals): with those as run-time
  // switches inside the unrolled epilogue loop the compiler emitted ~130 instructions per 8 outputs (register shuffles to
  // merge the paths, ~40 predicated-off residual instructions) instead of ~20.
  EPI_BF16_GELU = 6,
  EPI_BF16_RELU = 7,
  EPI_BF16_RESID = 8,
  // fp32 in-place residual update through TMA: out_f32[m, n] += (acc + bias) * gamma. The x tile is fetched into the
  // warp's swizzled staging buffer by TMA (prefetched one 32-column block ahead), updated in shared memory and stored back
  // by TMA, so the fp32 residual stream of the encoder (block.py:105-106) is read and written coalesced, under the GEMM's
  // tensor time, instead of by the LayerNorm kernel (which drops from 14 / 8 to 6 bytes per element).
  EPI_RESID_F32 = 9,
  // EPI_BF16 without activation / residuals, stored as IEEE fp16 instead of bf16: the per-tap contraction map of the fused
  // tail (output_conv2.0 applied at low resolution), whose consumer interpolates it with packed fp16 FMAs. Values are O(1);
  // fp16 keeps three more mantissa bits than bf16.
  EPI_F16 = 10,
  // conv3x3 + per-pixel LayerNorm over the channels + ReLU in one pass (input_projection of the guided head, dpt.py:153-159:
  // Conv2d, channels_first LayerNorm of dpt.py:37-61 with eps 1e-6 and biased variance, ReLU). Needs every channel of a
  // pixel in one accumulator tile: N <= 256, BN = 256. Each epilogue thread owns a whole 256-column row (warpgroup h drains
  // the tiles of accumulator stage h instead of half the columns of every tile), reads its row from tensor memory three
  // times (mean, variance, normalise) and stores bf16 through TMA. The statistics are taken from the fp32 accumulators,
  // not from a bf16-rounded conv output as the separate channel_ln_relu_kernel has to.
  // bias = conv bias, gamma = LayerNorm weight, aux = LayerNorm bias.
  EPI_BF16_CHLN = 11
};
__host__ __device__ constexpr bool epi_is_bf16(int e) {
  return e == EPI_BF16 || e == EPI_BF16_GELU || e == EPI_BF16_RELU || e == EPI_BF16_RESID || e == EPI_F16;
}
enum ActMode : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
enum AMode : int { A_LINEAR = 0, A_CONV3X3 = 1 };

struct GemmArgs {
  int M, N, K;          // logical problem (conv mode: M = B*H*W pixels, K = 9 * c_pad)
  int a_mode;
  int epi, act;
  // conv geometry (A_CONV3X3) -- also used by EPI_CONVT for the input grid
  int H, W, tiles_x, tiles_y, c_chunks;  // OUTPUT map size; c_chunks = c_pad / 64
  // A_CONV3X3 M tile = 2^lw x 2^lh pixels of 2^lb consecutive images (lw + lh + lb = 7; 16 x 8 x 1 by default): the shape is
  // picked per launch to minimise the padding of the (H, W, batch) extents -- 8 x 16 tiles pad a 37 x 37 map by 1.40x and a
  // 19 x 19 one by 2.13x, 2 x 2 pixels x 32 images pad them by 1.05x / 4 x 4 x 8 by 1.11x. tiles_x counts tiles of 2^lw * CG
  // pixels, tiles_b groups of 2^lb images. The accumulation order of an output element does not depend on the shape.
  int lw, lh, lb, tiles_b;
  int conv_taps;        // A_CONV3X3 (pixel-tile A operand): 9 = 3x3 conv, 1 = pointwise (k == s transposed conv, EPI_CONVT)
  int conv_stride;      // A_CONV3X3: 1, or 2 (resize_layers[3], dpt.py:102-107): input pixel = stride * output pixel + tap - 1
  // epilogue operands
  const float* bias;        // [N] (EPI_CONVT: [Cout]) or nullptr
  const float* gamma;       // [N] LayerScale, or nullptr
  float* out_f32;
  __nv_bfloat16* out_bf16;  // direct-store modes only (EPI_CONVT); TMA-store modes use tmap_c
  const __nv_bfloat16* resid1;
  const __nv_bfloat16* resid2;
  const float* aux;
  int ldo;                  // output row pitch in elements
  int P;                    // EPI_EMBED: patches per image
  int ks, cout;             // EPI_CONVT: kernel == stride, output channels
  int sigmoid;              // EPI_TAIL
  int has_relu_copy;        // EPI_BF16: also store relu(out) through tmap_c2
  int debug_timeline;       // bring-up builds (-DADA_BRINGUP) only: warp 4 lane 0 of CTA 0 stamps clock64 per epilogue phase
};

#ifndef ADA_RESID_BUFS
#define ADA_RESID_BUFS 2
#endif
#ifndef ADA_STG256
#define ADA_STG256 1
#endif
template <int BN, int CG, int EPI = 0>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBRows = BN / CG;                         // B rows (N) loaded by one CTA
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue staging: one 4 KB buffer per epilogue warp for the 256-wide tiles (two 64-column groups per warp and a long
  // main loop hide the TMA-store drain), two for narrower tiles (small-K, store-bound GEMMs: a single buffer made every
  // tile wait ~1 us for the previous store to release it)
  static constexpr int kStgBufs = (EPI == EPI_RESID_F32) ? ADA_RESID_BUFS : (BN == 256) ? ADA_STG256 : 2;  // RESID_F32: load-ahead (+ store-behind)
  static constexpr int kStagesFit = (232448 - 8 * kStgBufs * 4096 - 2 * 256 * 4 - 512) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = (BN * 2 < 32) ? 32 : BN * 2;  // 2 accumulator stages, power of two >= 32
  static constexpr int kStagingBytes = 8 * kStgBufs * 4096;      // 8 epilogue warps x kStgBufs x (32 rows x 128 B)
  static constexpr int kVecBytes = 2 * 256 * 4;                  // bias + gamma of the current N tile
  static constexpr int kBarBytes = 512;
  // no alignment slack: the kernel has no static shared memory, so the dynamic window starts 1 KB aligned (checked)
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kVecBytes + kBarBytes;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

// GELU with the exact-erf definition (nn.GELU default, mlp.py:30-41): gelu(x) = x * Phi(x). Phi is evaluated as
// sigmoid(x * q(x^2)) with q a degree-2 minimax fit (scipy, x in [-8, 8]) of logit(Phi(x)) / x:
//   max |approx - x * 0.5 * (1 + erf(x / sqrt 2))| = 2.5e-5 for the fit itself.
// sigmoid(z) = 0.5 + 0.5 tanh(z / 2) turns the two MUFU ops (ex2, rcp) into one (tanh.approx, relative error 2^-11, i.e.
// <= 2.4e-4 |x| on the result -- a quarter of a bf16 ulp of the stored activation), and the arithmetic runs on packed
// fp32 pairs: 5 issue slots per element instead of 9. The epilogue warps share their schedulers with the TMA and MMA
// issuing warps, so epilogue issue slots are what fc1 (K = 1024) is short of.
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x2) {
  float s0, s1;
  f2_unpack(f2_mul(x2, x2), s0, s1);
  const uint64_t s2 = f2_pack(fminf(s0, 50.0f), fminf(s1, 50.0f));  // q is monotone on [0, 52]; tanh is saturated beyond
  uint64_t q = f2_fma(f2_pack(-0.0007030335770476013f * 0.5f, -0.0007030335770476013f * 0.5f), s2,
                      f2_pack(0.07401129204396431f * 0.5f, 0.07401129204396431f * 0.5f));
  q = f2_fma(q, s2, f2_pack(1.5950157685717141f * 0.5f, 1.5950157685717141f * 0.5f));
  float z0, z1;
  f2_unpack(f2_mul(x2, q), z0, z1);  // x q(x^2) / 2
  const uint64_t hx = f2_mul(x2, f2_pack(0.5f, 0.5f));
  return f2_fma(hx, f2_pack(fast_tanh(z0), fast_tanh(z1)), hx);
}

__device__ __forceinline__ void add_bf16x8(float (&v)[8], const uint4 rr) {
  v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
  v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
}

// EPI is a compile-time parameter: with every epilogue inlined behind run-time switches the kernel was ~4500 SASS
// instructions and ncu showed the epilogue warps stalled on instruction fetch (stall_no_inst) on every tile.
template <int BN, int CG, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2,
                    const GemmArgs g) {
  using Cfg = GemmCfg<BN, CG, EPI>;
  static_assert(CG == 1 || CG == 2, "cta_group is 1 or 2");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  uint8_t* smem_gen = smem_raw;
  if (threadIdx.x == 0 && (smem_base & 1023u)) {  // swizzled tiles need a 1 KB aligned window
    g_dev_error[0] = 0xA10;
    __trap();
  }
  const uint32_t staging_base = smem_base + Cfg::kStages * Cfg::kStageBytes;  // 1 KB aligned (stage sizes are)
  const uint32_t vec_off = Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes;
  float* s_vec = reinterpret_cast<float*>(smem_gen + vec_off);  // [bias 256 | gamma 256]
  const uint32_t bar_base = smem_base + vec_off + Cfg::kVecBytes;
  // barrier layout: full[kStages], empty[kStages], tfull[2], tempty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Cfg::kStages + 4);
  // EPI_RESID_F32: one load barrier per epilogue warp and staging buffer
  auto xld_bar = [&](int w, int b) { return bar_base + 8u * (2 * Cfg::kStages + 6 + w * 3 + b); };  // up to 3 per warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;  // position in the CTA pair; rank 0 issues the MMAs

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), CG);   // one arrive.expect_tx per producing CTA (leader's barrier is the one used)
      mbar_init(empty_bar(s), 1);   // tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    if constexpr (EPI == EPI_RESID_F32) {
      for (int w = 0; w < 8; ++w)
        for (int b = 0; b < Cfg::kStgBufs; ++b) mbar_init(xld_bar(w, b), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), (EPI == EPI_BF16_CHLN ? 4 : 8) * CG);  // one arrive per epilogue warp (CHLN: per warp of the warpgroup that owns the stage) of every CTA in the pair
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_ptr_smem, Cfg::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  griddep_wait();  // (programmatic dependent launch) everything above overlapped the previous kernel's tail
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  // A tile = CG stacked 128-row sub-tiles (linear) or CG horizontally adjacent 8x16 pixel patches (conv; g.tiles_x
  // already counts 16*CG-pixel wide tiles). Work unit = CTA pair when CG == 2.
  const int tiles_n = (g.N + BN - 1) / BN;
  const int tiles_m = (g.a_mode == A_CONV3X3) ? g.tiles_b * g.tiles_x * g.tiles_y
                                              : (g.M + kBlockM * CG - 1) / (kBlockM * CG);
  const int num_tiles = tiles_m * tiles_n;
  const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;
  const int num_kb = (g.a_mode == A_CONV3X3) ? g.conv_taps * g.c_chunks : (g.K + kBlockK - 1) / kBlockK;
  const int tap_off = (g.conv_taps == 9) ? 1 : 0;  // 3x3: taps start one pixel up / left (the padding); pointwise: none

  // Register rebalancing: the kernel is compiled for 168 regs/thread (384 threads); the producer/MMA warpgroup gives
  // registers back and the two epilogue warpgroups take them (setmaxnreg must dominate each role's code for ptxas).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = unit; t < num_tiles; t += num_units) {
      // let the next kernel's CTAs be scheduled once every CTA of this grid is on its LAST tile: its prologue then runs
      // on SMs that drain early instead of competing with full waves (an immediate trigger cost ~1 % at batch 32)
      if (t + num_units >= num_tiles) griddep_launch_dependents();
      const int mt = t / tiles_n, nt = t % tiles_n;
      int img = 0, y0 = 0, x0 = 0;
      if (g.a_mode == A_CONV3X3) {
        const int per_img = g.tiles_x * g.tiles_y;
        img = (mt / per_img) << g.lb;
        const int r = mt % per_img;
        y0 = (r / g.tiles_x) << g.lh;
        x0 = ((r % g.tiles_x) * CG + static_cast<int>(cta_rank)) << g.lw;
      }
      const int m_row0 = (mt * CG + static_cast<int>(cta_rank)) * kBlockM;
      const int n_row0 = nt * BN + static_cast<int>(cta_rank) * Cfg::kBRows;
      // whole warp converged, one elected lane per instruction; the (tap, channel-chunk) walk of the implicit-GEMM conv is
      // kept in counters instead of a division per K block.
      int cc = 0, kx = 0, ky = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u, 0x100 + stage);
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kABytes;
        if constexpr (CG == 2) {
          const uint32_t fb = mapa_shared(full_bar(stage), 0);  // the leader's barrier collects both CTAs' bytes
          mbar_expect_tx_cluster_w(fb, Cfg::kStageBytes);
          if (g.a_mode == A_CONV3X3)
            tma_load_4d_cg2_w(sa, &tmap_a, fb, cc * kBlockK, x0 * g.conv_stride + kx - tap_off, y0 * g.conv_stride + ky - tap_off, img);
          else
            tma_load_2d_cg2_w(sa, &tmap_a, fb, kb * kBlockK, m_row0);
          tma_load_2d_cg2_w(sb, &tmap_b, fb, kb * kBlockK, n_row0);
        } else {
          mbar_expect_tx_w(full_bar(stage), Cfg::kStageBytes);
          if (g.a_mode == A_CONV3X3)
            tma_load_4d_w(sa, &tmap_a, full_bar(stage), cc * kBlockK, x0 * g.conv_stride + kx - tap_off, y0 * g.conv_stride + ky - tap_off, img);
          else
            tma_load_2d_w(sa, &tmap_a, full_bar(stage), kb * kBlockK, m_row0);
          tma_load_2d_w(sb, &tmap_b, full_bar(stage), kb * kBlockK, n_row0);
        }
        if (++cc == g.c_chunks) {
          cc = 0;
          if (++kx == 3) { kx = 0; ++ky; }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ------------------------------------------------------------ MMA issuer (leader CTA of the pair)
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM * CG, BN, 0, 0);
    const uint64_t da0 = make_smem_desc_sw128(smem_base, 16, 1024);
    const uint64_t db0 = make_smem_desc_sw128(smem_base + Cfg::kABytes, 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = unit; t < num_tiles; t += num_units) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 0x200 + acc);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase, 0x300 + stage);
        tc_fence_after();
        {  // whole warp converged, one elected lane per tcgen05 instruction (ptx.cuh: umma_bf16_ss_w)
          const uint32_t soff = static_cast<uint32_t>(stage) * (Cfg::kStageBytes >> 4);  // descriptor address units: 16 B
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = da0 + soff + k * (kUmmaK * 2 >> 4);
            const uint64_t db = db0 + soff + k * (kUmmaK * 2 >> 4);
            if constexpr (CG == 2)
              umma_bf16_ss_cg2_w(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else
              umma_bf16_ss_w(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CG == 2) {
            umma_commit_cg2_w(empty_bar(stage), 0x3);                 // frees this smem slot in BOTH CTAs
            if (kb == num_kb - 1) umma_commit_cg2_w(tfull_bar(acc), 0x3);  // accumulator complete, both epilogues
          } else {
            umma_commit_w(empty_bar(stage));                 // smem slot reusable once these MMAs retire
            if (kb == num_kb - 1) umma_commit_w(tfull_bar(acc));  // accumulator complete
          }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // epilogue warpgroup: takes every other 64-column group of the tile
    const int row = q * 32 + lane;     // accumulator row owned by this thread
    const int et = threadIdx.x - 128;  // 0..255 within the epilogue warps
    const uint32_t buf0 = staging_base + static_cast<uint32_t>(warp - 4) * (4096u * Cfg::kStgBufs);  // this warp's staging
    int sbuf = 0;
    const uint32_t st_row = static_cast<uint32_t>(lane) * 128u;
    const uint32_t st_sw = static_cast<uint32_t>(lane & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr bool tma_out = (epi_is_bf16(EPI) || EPI == EPI_SWIGLU);
    constexpr bool stage_vec = tma_out || EPI == EPI_RESID_F32;  // bias / gamma staged in shared memory per N tile
    uint32_t xld_phase = 0;  // EPI_RESID_F32: phase bit per staging buffer (bits 0..2)
    int xbuf = 0;            // EPI_RESID_F32: staging buffer of the next block to consume
#ifdef ADA_BRINGUP
    const bool tl = g.debug_timeline && blockIdx.x == 0 && warp == 4 && lane == 0;
    int tl_i = 0;
    auto stamp = [&](int k) {
      if (tl && tl_i < 60) g_dev_timeline[tl_i * 8 + k] = clock64();
    };
#else
    auto stamp = [](int) {};
#endif
    if constexpr (EPI == EPI_BF16_CHLN) {  // single N tile: the conv bias is staged once
      for (int i = et; i < BN; i += kEpiThreads) s_vec[i] = (g.bias != nullptr && i < g.N) ? __ldg(g.bias + i) : 0.0f;
      named_bar_sync(1, kEpiThreads);
    }
    for (int t = unit; t < num_tiles; t += num_units) {
      if constexpr (EPI == EPI_BF16_CHLN) {
        if (acc != half) {  // the other warpgroup owns this accumulator stage
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          continue;
        }
      }
      const int mt = t / tiles_n, nt = t % tiles_n;
      const int n0 = nt * BN;
      stamp(0);
      // ---- map accumulator row -> output row
      bool valid;
      long long orow;  // output row index (pixels / tokens)
      int m;           // logical A row (used by EMBED / CONVT)
      int img = 0, y0 = 0, x0 = 0;
      if (g.a_mode == A_CONV3X3) {
        const int per_img = g.tiles_x * g.tiles_y;
        img = (mt / per_img) << g.lb;
        const int r = mt % per_img;
        y0 = (r / g.tiles_x) << g.lh;
        x0 = ((r % g.tiles_x) * CG + static_cast<int>(cta_rank)) << g.lw;
        // accumulator row -> (image, y, x) inside the tile: x fastest, then y, then image -- the order the TMA box delivered
        const int x = x0 + (row & ((1 << g.lw) - 1));
        const int y = y0 + ((row >> g.lw) & ((1 << g.lh) - 1));
        const int ib = img + (row >> (g.lw + g.lh));
        valid = (y < g.H) && (x < g.W) && (static_cast<long long>(ib) * g.H * g.W < g.M);
        orow = (static_cast<long long>(ib) * g.H + y) * g.W + x;
        m = static_cast<int>(orow);
        // this warp's 32 rows form the box {2^lw, min(2^lh, 32 >> lw), 32 >> (lw + lh) or 1} of the output map, starting at:
        y0 += ((q * 32) >> g.lw) & ((1 << g.lh) - 1);
        img += (q * 32) >> (g.lw + g.lh);
      } else {
        m = (mt * CG + static_cast<int>(cta_rank)) * kBlockM + row;
        valid = m < g.M;
        orow = m;
      }
      if constexpr (EPI == EPI_EMBED) {
        const int b = m / g.P, p = m % g.P;
        orow = static_cast<long long>(b) * (g.P + 1) + 1 + p;
      }

      // ---- stage this N tile's bias / gamma in shared memory: barrier (previous tile's readers done) -> write -> barrier
      float* s_bias = s_vec;
      float* s_gamma = s_bias + 256;
      // (no bias and no gamma: the vectors are tile-invariant -- staged once for the first tile, then no barriers)
      if (stage_vec && (g.bias != nullptr || g.gamma != nullptr || t == unit)) {  // stage_vec is constexpr
        named_bar_sync(1, kEpiThreads);
        for (int i = et; i < BN; i += kEpiThreads) {
          const int n = n0 + i;
          const float bv = (g.bias != nullptr && n < g.N) ? __ldg(g.bias + n) : 0.0f;
          const float gv = (g.gamma != nullptr && n < g.N) ? __ldg(g.gamma + n) : 1.0f;
          s_bias[i] = (epi_is_bf16(EPI) || EPI == EPI_RESID_F32) ? bv * gv : bv;  // (acc + b) * gamma as fma(acc, gamma, b * gamma)
          s_gamma[i] = gv;
        }
        named_bar_sync(1, kEpiThreads);
      }

      // EPI_RESID_F32: this warp's blocks of the tile are 32 rows x 32 fp32 columns (one 4 KB staging buffer, 128-byte
      // rows); column groups cg = half, half + 2, ... of 64 columns, two blocks per group.
      const int ew = warp - 4;
      const int xrow0 = (mt * CG + static_cast<int>(cta_rank)) * kBlockM + q * 32;
      int nblk = 0;
      auto blk_col = [&](int k) { return n0 + (half + 2 * (k >> 1)) * 64 + (k & 1) * 32; };
      auto fetch = [&](int k, int b) {  // x block k -> staging buffer b (whole warp, elected lane)
        mbar_expect_tx_w(xld_bar(ew, b), 4096);
        tma_load_2d_w(buf0 + static_cast<uint32_t>(b) * 4096u, &tmap_c, xld_bar(ew, b), blk_col(k), xrow0);
      };
      if constexpr (EPI == EPI_RESID_F32) {
        for (int cg = half; cg < BN / 64; cg += 2)
          if (n0 + cg * 64 < g.N) nblk += 2;
        // the first block is fetched before the accumulator is complete: its latency hides under the main loop's tail
        if (nblk > 0) {
          bulk_wait_read_w<Cfg::kStgBufs - 2>();  // earlier stores that may still read this buffer have drained
          __syncwarp();
          fetch(0, xbuf);
        }
        // pull the NEXT tile's x blocks of this warp into L2 now: a whole tile period ahead of their use
        const int tn = t + num_units;
        if (tn < num_tiles) {
          const int mtn = tn / tiles_n, n0n = (tn % tiles_n) * BN;
          const int rown = (mtn * CG + static_cast<int>(cta_rank)) * kBlockM + q * 32;
          for (int cg = half; cg < BN / 64; cg += 2) {
            if (n0n + cg * 64 < g.N) {
              tma_prefetch_l2_2d_w(&tmap_c, n0n + cg * 64, rown);
              tma_prefetch_l2_2d_w(&tmap_c, n0n + cg * 64 + 32, rown);
            }
          }
        }
      }
      stamp(1);
      mbar_wait(tfull_bar(acc), acc_phase, 0x400 + acc);
      tc_fence_after();
      stamp(2);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);

      if constexpr (EPI == EPI_BF16_CHLN) {
        if constexpr (BN == 256) {
          const float inv_n = 1.0f / static_cast<float>(g.N);
          const int ngroups = (g.N + 63) >> 6;
          // pass 1: mean of (acc + conv bias) over the N channels of this pixel
          float sum = 0.f;
#pragma unroll 1
          for (int c32 = 0; c32 < 2 * ngroups; ++c32) {
            uint32_t r[32];
            tmem_ld32(t_addr + c32 * 32, r);
            tmem_ld_wait();
            const float* bb = s_bias + c32 * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (c32 * 32 + j < g.N) {  // N % 8 == 0
#pragma unroll
                for (int k = 0; k < 8; ++k) sum += __uint_as_float(r[j + k]) + bb[j + k];
              }
          }
          const float mean = sum * inv_n;
          // pass 2: biased variance (two-pass, as dpt.py:56-58)
          float sq = 0.f;
#pragma unroll 1
          for (int c32 = 0; c32 < 2 * ngroups; ++c32) {
            uint32_t r[32];
            tmem_ld32(t_addr + c32 * 32, r);
            tmem_ld_wait();
            const float* bb = s_bias + c32 * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (c32 * 32 + j < g.N) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float d = __uint_as_float(r[j + k]) + bb[j + k] - mean;
                  sq = fmaf(d, d, sq);
                }
              }
          }
          const float rstd = 1.0f / sqrtf(sq * inv_n + 1e-6f);
          // pass 3: normalise, affine, ReLU, bf16, TMA store (columns past N are clipped by the tensor map)
#pragma unroll 1
          for (int cg = 0; cg < ngroups; ++cg) {
            const int oc = cg * 64;
            uint32_t pk[32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t r[32];
              tmem_ld32(t_addr + oc + h * 32, r);
              tmem_ld_wait();
              const float* bb = s_bias + oc + h * 32;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int col = oc + h * 32 + j;
                float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = w4;
                if (col < g.N) {
                  w4 = __ldg(reinterpret_cast<const float4*>(g.gamma + col));
                  b4 = __ldg(reinterpret_cast<const float4*>(g.aux + col));
                }
                const float y0 = fmaxf(fmaf((__uint_as_float(r[j]) + bb[j] - mean) * rstd, w4.x, b4.x), 0.f);
                const float y1 = fmaxf(fmaf((__uint_as_float(r[j + 1]) + bb[j + 1] - mean) * rstd, w4.y, b4.y), 0.f);
                const float y2 = fmaxf(fmaf((__uint_as_float(r[j + 2]) + bb[j + 2] - mean) * rstd, w4.z, b4.z), 0.f);
                const float y3 = fmaxf(fmaf((__uint_as_float(r[j + 3]) + bb[j + 3] - mean) * rstd, w4.w, b4.w), 0.f);
                pk[h * 16 + (j >> 1)] = pack_bf16x2(y0, y1);
                pk[h * 16 + (j >> 1) + 1] = pack_bf16x2(y2, y3);
              }
            }
            const uint32_t buf = buf0 + static_cast<uint32_t>(sbuf) * 4096u;
            bulk_wait_read_w<Cfg::kStgBufs - 1>();  // (elected lane) the store that last used this buffer has drained it
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c)
              st_shared_v4(buf + st_row + ((static_cast<uint32_t>(c) ^ st_sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2],
                           pk[4 * c + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (g.a_mode == A_CONV3X3)
              tma_store_4d_commit_w(&tmap_c, buf, oc, x0, y0, img);
            else
              tma_store_2d_commit_w(&tmap_c, buf, oc, (mt * CG + static_cast<int>(cta_rank)) * kBlockM + q * 32);
            if constexpr (Cfg::kStgBufs == 2) sbuf ^= 1;
          }
        }
      } else if constexpr (tma_out) {
        if constexpr (BN >= 64) {
          constexpr int out_cols = (EPI == EPI_SWIGLU) ? BN / 2 : BN;  // output columns produced by this tile
          const int on0 = (EPI == EPI_SWIGLU) ? (n0 >> 1) : n0;
          const int n_out = (EPI == EPI_SWIGLU) ? (g.N >> 1) : g.N;
#pragma unroll 1
          for (int cg = half; cg < out_cols / 64; cg += 2) {
            const int oc = on0 + cg * 64;  // first output column of this 64-wide group
            if (oc >= n_out) break;
            uint32_t pk[32];               // 64 bf16 outputs of this thread's row
            if constexpr (EPI == EPI_SWIGLU) {
              // 128 interleaved accumulator columns: [x1 32 | x2 32 | x1 32 | x2 32]
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t a[32], b[32];
                tmem_ld32(t_addr + cg * 128 + h * 64, a);
                tmem_ld32(t_addr + cg * 128 + h * 64 + 32, b);
                tmem_ld_wait();
                const float* bb = s_bias + cg * 128 + h * 64;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  // silu(x1) * x2 on packed pairs; silu(x) = x/2 + x/2 * tanh(x/2): one MUFU instead of ex2 + rcp
                  const uint64_t x1 = f2_add(f2_pack(__uint_as_float(a[j]), __uint_as_float(a[j + 1])),
                                             *reinterpret_cast<const uint64_t*>(bb + j));
                  const uint64_t x2 = f2_add(f2_pack(__uint_as_float(b[j]), __uint_as_float(b[j + 1])),
                                             *reinterpret_cast<const uint64_t*>(bb + 32 + j));
                  const uint64_t hx = f2_mul(x1, f2_pack(0.5f, 0.5f));
                  float h0, h1;
                  f2_unpack(hx, h0, h1);
                  float ha, hb;
                  f2_unpack(f2_mul(f2_fma(hx, f2_pack(fast_tanh(h0), fast_tanh(h1)), hx), x2), ha, hb);
                  pk[h * 16 + (j >> 1)] = pack_bf16x2(ha, hb);
                }
              }
            } else {
              // residual rows of this 64-column group: all 16-byte loads are issued before the accumulator is touched. A
              // thread reads its own output row, so nothing coalesces and every load is a DRAM / L2 round trip; issued
              // one pair at a time next to their use (ncu: the epilogue warps sat on long-scoreboard stalls at the first
              // use of each loaded value) the epilogue of a 2-residual + ReLU-copy conv took longer than the K loop of the
              // next tile: tensor pipe 54 % (877 us at 148^2 x 256 channels, batch 32) against 94 % (507 us) without residuals.
              uint4 rz1[8], rz2[8];
              if constexpr (EPI == EPI_BF16_RESID) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  rz1[c] = make_uint4(0u, 0u, 0u, 0u);
                  rz2[c] = make_uint4(0u, 0u, 0u, 0u);
                }
                if (valid) {
                  const long long off0 = orow * g.ldo + oc;
                  if (g.resid1) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                      if (oc + c * 8 < g.N) rz1[c] = *reinterpret_cast<const uint4*>(g.resid1 + off0 + c * 8);
                  }
                  if (g.resid2) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                      if (oc + c * 8 < g.N) rz2[c] = *reinterpret_cast<const uint4*>(g.resid2 + off0 + c * 8);
                  }
                }
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(t_addr + cg * 64 + h * 32, r);
                tmem_ld_wait();
                const float* bb = s_bias + cg * 64 + h * 32;
                const float* gg = s_gamma + cg * 64 + h * 32;
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                  // (acc + bias) * gamma as one packed FMA per pair: the staged bias is already multiplied by gamma and
                  // the staged gamma is 1 when there is none (no run-time switch inside the unrolled loop)
                  uint64_t v2[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint64_t a2 = f2_pack(__uint_as_float(r[gi * 8 + 2 * j]), __uint_as_float(r[gi * 8 + 2 * j + 1]));
                    v2[j] = f2_fma(a2, *reinterpret_cast<const uint64_t*>(gg + gi * 8 + 2 * j),
                                   *reinterpret_cast<const uint64_t*>(bb + gi * 8 + 2 * j));
                  }
                  if constexpr (EPI == EPI_BF16_GELU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v2[j] = gelu_erf2(v2[j]);
                  }
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 4; ++j) f2_unpack(v2[j], v[2 * j], v[2 * j + 1]);
                  if constexpr (EPI == EPI_BF16_RELU) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
                  }
                  if constexpr (EPI == EPI_BF16_RESID) {
                    if (g.act == ACT_RELU) {  // (only the operator-level ABI combines ReLU with residual adds)
#pragma unroll
                      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
                    }
                    const int n = oc + h * 32 + gi * 8;
                    if (valid && n < g.N) {
                      if (g.resid1) add_bf16x8(v, rz1[h * 4 + gi]);
                      if (g.resid2) add_bf16x8(v, rz2[h * 4 + gi]);
                    }
                  }
#pragma unroll
                  for (int j = 0; j < 8; j += 2)
                    pk[h * 16 + gi * 4 + (j >> 1)] = (EPI == EPI_F16) ? pack_f16x2(v[j], v[j + 1]) : pack_bf16x2(v[j], v[j + 1]);
                }
              }
            }
            // ---- registers -> swizzled staging buffer -> TMA store
            stamp(3);
            const int passes = g.has_relu_copy ? 2 : 1;
            for (int pass = 0; pass < passes; ++pass) {
              const uint32_t buf = buf0 + static_cast<uint32_t>(sbuf) * 4096u;
              bulk_wait_read_w<Cfg::kStgBufs - 1>();  // (elected lane) the store that last used this buffer has drained it
              __syncwarp();
              stamp(4);
              if (pass == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {  // relu on packed bf16 pairs: clear negatives (sign bit set)
                  uint32_t w = pk[i];
                  if (w & 0x00008000u) w &= 0xFFFF0000u;
                  if (w & 0x80000000u) w &= 0x0000FFFFu;
                  pk[i] = w;
                }
              }
#pragma unroll
              for (int c = 0; c < 8; ++c)
                st_shared_v4(buf + st_row + ((static_cast<uint32_t>(c) ^ st_sw) << 4), pk[4 * c], pk[4 * c + 1],
                             pk[4 * c + 2], pk[4 * c + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              stamp(5);
              {  // whole warp converged, elected lane issues store + commit (no uniform-register waterfall)
                const CUtensorMap* tm = (pass == 0) ? &tmap_c : &tmap_c2;
                if (g.a_mode == A_CONV3X3)
                  tma_store_4d_commit_w(tm, buf, oc, x0, y0, img);
                else
                  tma_store_2d_commit_w(tm, buf, oc, (mt * CG + static_cast<int>(cta_rank)) * kBlockM + q * 32);
              }
              if constexpr (Cfg::kStgBufs == 2) sbuf ^= 1;
            }
          }
        }
      } else if constexpr (EPI == EPI_RESID_F32) {
        if constexpr (BN >= 64) {
#pragma unroll 1
          for (int k = 0; k < nblk; ++k) {
            const int b = xbuf;
            const int bn_ = (b == Cfg::kStgBufs - 1) ? 0 : b + 1;
            if (k + 1 < nblk) {  // prefetch the next block once the store that last read its buffer has drained
              bulk_wait_read_w<Cfg::kStgBufs - 2>();
              __syncwarp();
              fetch(k + 1, bn_);
            }
            const int c_local = (half + 2 * (k >> 1)) * 64 + (k & 1) * 32;  // first accumulator column of this block
            uint32_t r[32];
            tmem_ld32(t_addr + c_local, r);
            tmem_ld_wait();
            mbar_wait(xld_bar(ew, b), (xld_phase >> b) & 1u, 0x480 + b);
            xld_phase ^= 1u << b;
            const uint32_t buf = buf0 + static_cast<uint32_t>(b) * 4096u;
            const float* bb = s_bias + c_local;
            const float* gg = s_gamma + c_local;
#pragma unroll
            for (int c = 0; c < 8; ++c) {  // 4 fp32 per 16-byte chunk, chunk index XOR (row & 7) = the 128-byte swizzle
              const uint32_t addr = buf + st_row + ((static_cast<uint32_t>(c) ^ st_sw) << 4);
              uint64_t x01, x23;
              asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x01), "=l"(x23) : "r"(addr));
              const uint64_t v01 = f2_fma(f2_pack(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1])),
                                          *reinterpret_cast<const uint64_t*>(gg + 4 * c), *reinterpret_cast<const uint64_t*>(bb + 4 * c));
              const uint64_t v23 = f2_fma(f2_pack(__uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3])),
                                          *reinterpret_cast<const uint64_t*>(gg + 4 * c + 2),
                                          *reinterpret_cast<const uint64_t*>(bb + 4 * c + 2));
              x01 = f2_add(x01, v01);
              x23 = f2_add(x23, v23);
              asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(x01), "l"(x23) : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            tma_store_2d_commit_w(&tmap_c, buf, blk_col(k), xrow0);  // clips rows >= M and columns >= N
            xbuf = bn_;
          }
        }
      } else if constexpr (EPI == EPI_TAIL) {
        if constexpr (BN == 32) if (half == 0) {
          uint32_t r[32];
          tmem_ld32(t_addr, r);
          tmem_ld_wait();
          float s = __ldg(g.aux + 32);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(r[j]) + __ldg(g.bias + j);
            v = fmaxf(v, 0.0f);
            s = fmaf(v, __ldg(g.aux + j), s);
          }
          if (g.sigmoid == 1) s = 1.0f / (1.0f + __expf(-s));
          else if (g.sigmoid == 2) s = fmaxf(s, 0.f);
          if (valid) g.out_f32[orow] = s;
        }
      } else {
        // EPI_CONVT on pixel tiles (A operand fetched as 8x16 pixel patches, Cout % 64 == 0): the pixel shuffle of a k == s
        // transposed conv is a 5-D tensor (kx*Cout + co, x, ky, y, b); a 64-column group of the GEMM output is one (ky, kx) and
        // 64 consecutive channels, and the 32 rows of a warp are two 16-pixel rows of the patch -- one TMA box, the parts past
        // the image clipped. (TMA stores take no negative coordinates -- illegal instruction, tools/micro/tma5d_test.cu -- so a
        // linear M tiling, whose warps wrap around image rows, cannot be expressed as boxes.) The per-thread scatter this
        // replaces wrote 16 bytes per lane to 32 different cache lines (262 - 485 TFLOP/s on the two resize layers).
        bool convt_tma = false;
        if constexpr (EPI == EPI_CONVT && BN >= 64) convt_tma = (g.a_mode == A_CONV3X3) && (g.cout % 64) == 0;
        if (convt_tma) {
          if constexpr (EPI == EPI_CONVT && BN >= 64) {
#pragma unroll 1
            for (int cg = half; cg < BN / 64; cg += 2) {
              const int oc = n0 + cg * 64;
              if (oc >= g.N) break;
              const int kk = oc / g.cout, co0 = oc - kk * g.cout;
              const int ky = kk / g.ks, kx = kk - ky * g.ks;
              uint32_t pk[32];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(t_addr + cg * 64 + h * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (g.bias) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + co0 + h * 32 + j));
                  pk[h * 16 + (j >> 1)] = pack_bf16x2(__uint_as_float(r[j]) + b4.x, __uint_as_float(r[j + 1]) + b4.y);
                  pk[h * 16 + (j >> 1) + 1] = pack_bf16x2(__uint_as_float(r[j + 2]) + b4.z, __uint_as_float(r[j + 3]) + b4.w);
                }
              }
              const uint32_t buf = buf0 + static_cast<uint32_t>(sbuf) * 4096u;
              bulk_wait_read_w<Cfg::kStgBufs - 1>();
              __syncwarp();
#pragma unroll
              for (int c = 0; c < 8; ++c)
                st_shared_v4(buf + st_row + ((static_cast<uint32_t>(c) ^ st_sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2],
                             pk[4 * c + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              tma_store_5d_w(&tmap_c, buf, kx * g.cout + co0, x0, ky, y0, img);
              bulk_commit_w();
              if constexpr (Cfg::kStgBufs == 2) sbuf ^= 1;
            }
          }
        } else {
        // EPI_EMBED / EPI_CONVT: direct per-thread stores (one GEMM each per forward; row remap / pixel-shuffle scatter)
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld32(t_addr + c * 32, r);
          tmem_ld_wait();
          const int nb = n0 + c * 32;
          if (!valid || nb >= g.N) continue;
          if constexpr (EPI == EPI_EMBED) {
            const int p = m % g.P;
            const float* ax = g.aux + static_cast<long long>(p) * g.N + nb;
            float* dst = g.out_f32 + orow * g.ldo + nb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (nb + j < g.N) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(ax + j));
                float4 o;
                o.x = __uint_as_float(r[j]) + a4.x;
                o.y = __uint_as_float(r[j + 1]) + a4.y;
                o.z = __uint_as_float(r[j + 2]) + a4.z;
                o.w = __uint_as_float(r[j + 3]) + a4.w;
                *reinterpret_cast<float4*>(dst + j) = o;
              }
            }
          } else {
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              const int n = nb + gi * 8;
              if (n >= g.N) break;
              const int kk = n / g.cout, co = n % g.cout;
              const int ky = kk / g.ks, kx = kk % g.ks;
              const int hw = g.H * g.W;
              const int b = m / hw, rem = m % hw;
              const int y = rem / g.W, x = rem % g.W;
              const long long off =
                  ((static_cast<long long>(b) * g.H * g.ks + y * g.ks + ky) * (g.W * g.ks) + x * g.ks + kx) * g.cout + co;
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[gi * 8 + j]);
              if (g.bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + co));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + co + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              *reinterpret_cast<uint4*>(g.out_bf16 + off) = make_uint4(
                  pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
          }
        }
        }
      }
      // accumulator stage drained -> hand it back to the MMA warp
      stamp(6);
#ifdef ADA_BRINGUP
      ++tl_i;
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2)
          mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));  // the leader's MMA warp owns the accumulator hand-off
        else
          mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    bulk_wait_w<0>();  // (same elected lane) all TMA stores of this warp have completed before the CTA retires
    __syncwarp();
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();  // the peer may still read our smem / TMEM until here
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace ada
