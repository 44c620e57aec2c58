// Fused flash-style multi-head attention for sm_100a, head_dim 64 (every DINOv2 size), no mask, no dropout.
// Replaces  softmax((q * d^-0.5) k^T) v  of attention.py:53-59 (== xformers memory_efficient_attention, attention.py:74-77)
// without materialising the N x N score matrix.
//
// One CTA = 128 queries of one (image, head); Q/K/V are read in place from the [B, N, 3, H, 64] QKV GEMM output
// through one 3-D TMA map (out-of-range tokens are zero-filled by the hardware).
//   warp 0     : TMA producer (Q once, then a 2-stage K ring and a 2-stage V ring)
//   warp 1     : tcgen05 issuer, the whole warp converged with one elected lane per instruction (ptx.cuh umma_*_w).
//                                 S = Q K^T  (128x128x64, both operands K-major)        -> TMEM cols [0,128)
//                                 O += P V   (128x64x128, P from TMEM, V MN-major smem)  -> TMEM cols [192,256)
//   warps 2..9 : softmax, EIGHT warps: every query row (= TMEM lane) is shared by two threads that each own 64 of the
//                128 score columns (warps w and w+4 sit on the same TMEM lane quarter). The pair exchanges its partial
//                row maximum through shared memory; everything else is thread-private. ncu on the 4-warp version showed
//                the softmax warps busy 83% of the time at 0.25 IPC each (fixed-latency dependency stalls) with the
//                tensor pipe at 30% and MUFU at 61%: the kernel needed more warps per scheduler, not fewer instructions.
//                P is written back to TMEM (cols [128,192), bf16 pairs) with tcgen05.st and consumed by the P V MMA
//                straight from there; S is released to the issuer as soon as it sits in registers.
// The running output stays in TMEM across KV tiles. Rows are kept relative to a *stale* maximum: the accumulator is only
// rescaled (tcgen05.ld -> scale -> tcgen05.st) when some row's maximum grows by more than 2^8, which happens in the first
// tile or two; p may then reach 256, harmless in fp32/bf16, and O / l is exact either way.
// Two CTAs fit per SM (80 KB smem, 256 TMEM columns, 96 registers x 320 threads each).
// Variants tried and measured on the same shape (B=32, N=1370, 16 heads; numbers and timelines in profiles/README.md):
// 64-key tiles with S and P double-buffered (0.42 ms), two query tiles per CTA sharing K/V (0.49 ms), fp16-pair
// exponentials (ex2.approx.f16x2 is issued as two MUFU.EX2.F16, no gain), two softmax groups ping-ponging 64-key tiles
// with a shared running maximum (0.371 ms) or as two independent streams merged in the epilogue (0.384 ms), against
// 0.368 ms for this kernel. Issue slots ~55 %, MUFU 40-64 %, tensor pipe ~31 %: no single pipe is saturated; the two
// resident CTAs' fixed-latency phases (row max + pair exchange, TMEM ld/st, barrier round trips) only partly overlap.
#pragma once
#include "ptx.cuh"

namespace ada {

constexpr int kAttThreads = 384;   // warps 0..3: producer, score issuer, P V issuer, idle; warps 4..11: softmax
constexpr int kAttSoftmaxThreads = 256;
constexpr int kAttQ = 128, kAttKV = 128, kAttD = 64;
constexpr int kAttSmemBytes = 16384 /*Q*/ + 2 * 16384 /*K*/ + 2 * 16384 /*V*/ + 3072 /*pair exchange*/ + 256 /*barriers*/;
constexpr int kAttTmemCols = 256;
constexpr float kAttRescaleLog2 = 8.0f;  // rescale O only when a row max grows by more than 2^8

struct AttArgs {
  int B, N, heads, D;        // D = heads * 64
  float scale_log2e;         // d^-0.5 * log2(e)
};

// VARIANT is a measurement knob of bring-up builds (-DADA_BRINGUP, env ADA_ATT_VARIANT; the product library instantiates
// variant 0 only): 0 = product (6 of every 16 exponential pairs, evenly spaced, on the
// FMA pipe), 1 = every exponential on MUFU, 3 = 6/16 clustered, 4 = 8/16, 5 = 4/16, 2 = exponentials replaced by a copy
// (timing skeleton only, wrong results), 10 = clock64 timeline of one softmax thread and of the issuer thread.
template <bool B>
struct IntTag {
  static constexpr bool value = B;
};

template <int VARIANT>
__global__ void __launch_bounds__(kAttThreads, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                         const AttArgs a) {
  extern __shared__ __align__(1024) uint8_t att_smem[];
  const uint32_t sbase = smem_u32(att_smem);
  const uint32_t sQ = sbase, sK = sbase + 16384, sV = sbase + 49152;
  float* xch = reinterpret_cast<float*>(att_smem + 81920);  // [3][2][128]: row-max exchange (2 parities) + row sums
  const uint32_t bar = sbase + 81920 + 3072;
  const uint32_t q_full = bar, s_full = bar + 8, p_full = bar + 16, o_full = bar + 24, s_free = bar + 32;
  auto k_full = [&](int s) { return bar + 40 + 8u * s; };
  auto k_empty = [&](int s) { return bar + 56 + 8u * s; };
  auto v_full = [&](int s) { return bar + 72 + 8u * s; };
  auto v_empty = [&](int s) { return bar + 88 + 8u * s; };
  const uint32_t tmem_ptr_smem = bar + 104;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttQ, head = blockIdx.y, img = blockIdx.z;
  const int num_kv = (a.N + kAttKV - 1) / kAttKV;

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {  // the swizzled layouts below assume a 1 KB aligned window
      g_dev_error[0] = 0xA11;
      __trap();
    }
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, kAttSoftmaxThreads);
    mbar_init(o_full, 1);
    mbar_init(s_free, kAttSoftmaxThreads);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, kAttTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();
  griddep_launch_dependents();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;

  // registers: compiled for 80 per thread (two CTAs of 12 warps per SM); the control warpgroup keeps 48, the two softmax
  // warpgroups take 96 (setmaxnreg)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
  if (warp == 0) {
    // ------------------------------------------------------------------ producer (whole warp, elected lane per instruction)
    mbar_expect_tx_w(q_full, 16384);
    tma_load_3d_w(sQ, &tmap_qkv, q_full, head * kAttD, q0, img);
    for (int j = 0; j < num_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(k_empty(s), ph ^ 1u, 0x500 + s);
      mbar_expect_tx_w(k_full(s), 16384);
      tma_load_3d_w(sK + s * 16384, &tmap_qkv, k_full(s), a.D + head * kAttD, j * kAttKV, img);
      mbar_wait(v_empty(s), ph ^ 1u, 0x510 + s);
      mbar_expect_tx_w(v_full(s), 16384);
      tma_load_3d_w(sV + s * 16384, &tmap_qkv, v_full(s), 2 * a.D + head * kAttD, j * kAttKV, img);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint64_t dq = make_smem_desc_sw128(sQ, 16, 1024);
    const uint64_t dk0 = make_smem_desc_sw128(sK, 16, 1024);
    auto issue_s = [&](int j) {
      const int s = j & 1;
      mbar_wait(k_full(s), (j >> 1) & 1, 0x520 + s);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss_w(tS, dq + 2 * k, dk0 + s * 1024 + 2 * k, idesc_s, k > 0 ? 1u : 0u);
      umma_commit_w(k_empty(s));
      umma_commit_w(s_full);
    };
#ifdef ADA_BRINGUP
    const bool tli = (VARIANT == 10) && lane == 0 && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 7;
    auto istamp = [&](int j, int k) {
      if (tli) g_dev_timeline[128 + j * 8 + k] = clock64();
    };
#else
    auto istamp = [](int, int) {};
#endif
    mbar_wait(q_full, 0, 0x530);
    issue_s(0);
    for (int j = 0; j + 1 < num_kv; ++j) {
      named_bar_sync(6, kAttSoftmaxThreads + 32);  // softmax has pulled S(j) into registers -> S(j+1) may overwrite it
      issue_s(j + 1);
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ P V issuer, on another scheduler than the score
    // issuer: a tcgen05.mma shares its scheduler's MIO queue with the MUFU instructions of the softmax warps living there
    // (~100 cycles per MMA to get through it); with the 12 MMAs of a tile issued by one warp -- and the issuers of both
    // resident CTAs on the same scheduler -- MMA issue was what bounded the kernel (clock64 timelines, profiles/README.md).
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B = V is MN-major (d contiguous)
    const uint64_t dv0 = make_smem_desc_sw128(sV, 0, 1024);
    for (int j = 0; j < num_kv; ++j) {
      const int s = j & 1;
      named_bar_sync(7, kAttSoftmaxThreads + 32);  // P(j) in TMEM, O rescaled if needed
      mbar_wait(v_full(s), (j >> 1) & 1, 0x550 + s);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)  // 16 keys per MMA = 8 packed TMEM columns of P
        umma_bf16_ts_w(tO, tP + kk * 8, dv0 + s * 1024 + kk * 128, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
      umma_commit_w(v_empty(s));
      umma_commit_w(o_full);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    // ------------------------------------------------------------------ softmax / output warps
    const int qd = warp & 3;               // TMEM lane quarter
    const int half = (warp - 4) >> 2;      // which 64 score columns of the row this thread owns
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float c = a.scale_log2e;
    float m_used = -INFINITY, l_part = 0.f;

#ifdef ADA_BRINGUP
    const bool tl = (VARIANT == 10) && threadIdx.x == 128 && blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 7;
    auto stamp = [&](int j, int k) {
      if (tl) g_dev_timeline[j * 8 + k] = clock64();
    };
#else
    auto stamp = [](int, int) {};
#endif
    // One KV tile. MASKED is a compile-time tag: only the ragged last tile carries the 64 compare+select pairs that
    // overwrite the scores of keys past N (zero-filled by TMA) with -inf. As a run-time `if` inside a single loop body the
    // compiler if-converted them into ~130 always-executed instructions per tile (of ~450).
    auto tile = [&](const int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int kv_valid = min(kAttKV, a.N - j * kAttKV) - half * 64;  // valid keys inside this thread's 64 columns
      stamp(j, 0);
      mbar_wait(s_full, j & 1, 0x560);
      stamp(j, 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tS + lane_off + half * 64, s0);
      tmem_ld32(tS + lane_off + half * 64 + 32, s1);
      tmem_ld_wait();
      stamp(j, 2);
      tc_fence_before();
      // this thread's part of S(j) now lives in registers. Hardware named barriers (arrive here, sync in the issuer warp)
      // hand off in tens of cycles; the mbarrier round trip they replace cost ~350 cycles per hop and made the kernel
      // synchronisation-latency bound (skeleton 0.23 ms of 0.39 ms with all math removed).
      if (j + 1 < num_kv) named_bar_arrive(6, kAttSoftmaxThreads + 32);
      if constexpr (MASKED) {  // ragged last tile: keys past N are zero-filled by TMA -> mask them out
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) s0[i] = 0xff800000u;
          if (32 + i >= kv_valid) s1[i] = 0xff800000u;
        }
      }
      float tm[4];  // FMNMX3: two scores per instruction, four independent chains of 16 scores
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t* v = (k < 2) ? (s0 + 16 * k) : (s1 + 16 * (k - 2));
        float t = fmax3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
#pragma unroll
        for (int i = 3; i < 15; i += 2) t = fmax3(t, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        tm[k] = fmaxf(t, __uint_as_float(v[15]));
      }
      float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
      // ---- the two owners of a row combine their partial maxima (double-buffered by tile parity)
      float* xm = xch + (j & 1) * 256;
      xm[half * 128 + row] = tmax;
      named_bar_sync(1 + qd, 64);
      tmax = fmaxf(tmax, xm[(half ^ 1) * 128 + row]);
      stamp(j, 3);

      // ---- running (stale) maximum: both owners take the same decision from the same combined maximum
      float sc = 1.0f;
      bool rescale = false;
      if (j == 0) {
        m_used = tmax;
      } else {
        const bool grow = (tmax - m_used) * c > kAttRescaleLog2;
        rescale = __any_sync(0xffffffffu, grow);  // rare (first tile or two); identical in both warps of the pair
        if (rescale) {
          const float m_new = fmaxf(m_used, tmax);
          sc = fast_exp2((m_used - m_new) * c);
          m_used = m_new;
          l_part *= sc;
        }
      }
      const float mc = m_used * c;
      // Scale-and-shift and the row sums run as packed pairs (FFMA2 / FADD2: 182 M instead of 267 M warp instructions per
      // launch at B=32); kEmuMask picks the pairs (of the 16 in a 32-column half) whose exponential is evaluated on the
      // FMA pipe instead of MUFU. Measured stand-alone (B=32, N=1370): none 0.379 ms, 4/16 0.371, 6/16 0.381, 8/16 0.410;
      // inside the model (1.55 GHz, power capped), ms per step over the 24 launches: none 12.3, 4/16 (every 4th pair) 11.4,
      // 6/16 clustered 11.8, 6/16 evenly spaced (the product mask) 11.15, 8/16 12.8. See profiles/README.md.
      constexpr uint32_t kEmuMask = (VARIANT == 1) ? 0u : (VARIANT == 3) ? 0x5252u : (VARIANT == 4) ? 0x5555u
                                  : (VARIANT == 5) ? 0x1111u : 0x9249u;
      const uint64_t c2 = f2_pack(c, c), nmc2 = f2_pack(-mc, -mc);
      uint64_t rs2[4] = {0ull, 0ull, 0ull, 0ull};  // independent packed row-sum chains
      uint32_t pk[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(h ? s1[i] : s0[i]), __uint_as_float(h ? s1[i + 1] : s0[i + 1])),
                                     c2, nmc2);
          float p0, p1;
          if constexpr (VARIANT == 2) {
            f2_unpack(x2, p0, p1);
          } else {
            if ((kEmuMask >> (i >> 1)) & 1u) {
              exp2_fma2(x2, p0, p1);
            } else {
              float x0, x1;
              f2_unpack(x2, x0, x1);
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
          }
          rs2[(i >> 1) & 3] = f2_add(rs2[(i >> 1) & 3], f2_pack(p0, p1));
          pk[h * 16 + (i >> 1)] = pack_bf16x2(p0, p1);
        }
      }
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(f2_add(rs2[0], rs2[1]), f2_add(rs2[2], rs2[3])), a0, a1);
        b0 = a0 + a1;
        (void)b1;
        l_part += b0;
      }
      stamp(j, 4);
      if (j > 0) {  // P(j-1) V(j-1) must have retired before P is overwritten / O may be rescaled
        mbar_wait(o_full, (j - 1) & 1, 0x570);
        tc_fence_after();
        if (rescale) {  // bring this warp's share (32 of the 64 columns) of its 32 accumulator rows to the new maxima
#pragma unroll 1
          for (int h = 0; h < 4; ++h) {
            uint32_t r[8];
            tmem_ld8(tO + lane_off + half * 32 + h * 8, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * sc);
            tmem_st8(tO + lane_off + half * 32 + h * 8, r);
          }
        }
      }
      stamp(j, 5);
      tmem_st32(tP + lane_off + half * 32, pk);
      tmem_st_wait();
      stamp(j, 6);
      tc_fence_before();
      named_bar_arrive(7, kAttSoftmaxThreads + 32);
    };
    const bool ragged = (a.N % kAttKV) != 0;
#pragma unroll 1
    for (int j = 0; j < num_kv - 1; ++j) tile(j, IntTag<false>{});
    if (ragged)
      tile(num_kv - 1, IntTag<true>{});
    else
      tile(num_kv - 1, IntTag<false>{});
    // ---- epilogue: O / l -> bf16 -> swizzled smem (the Q tile is dead once the last S MMA has retired) -> one TMA
    //      store per CTA (row-per-thread global stores touch 32 cache lines per warp instruction).
    float* xl = xch + 512;
    xl[half * 128 + row] = l_part;
    mbar_wait(o_full, (num_kv - 1) & 1, 0x580);
    tc_fence_after();
    named_bar_sync(1 + qd, 64);
    const float inv = 1.0f / (l_part + xl[(half ^ 1) * 128 + row]);
    const uint32_t o_row = sQ + static_cast<uint32_t>(row) * 128u;
    const uint32_t o_sw = static_cast<uint32_t>(row & 7);
    {
      uint32_t r[32];
      tmem_ld32(tO + lane_off + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t chunk = static_cast<uint32_t>(half * 4 + i);
        st_shared_v4(o_row + ((chunk ^ o_sw) << 4),
                     pack_bf16x2(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv),
                     pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv),
                     pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv),
                     pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv));
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(5, kAttSoftmaxThreads);
    if (threadIdx.x == 128) {  // first softmax thread; rows past N are clipped by the tensor map
      tma_store_3d(&tmap_out, sQ, head * kAttD, q0, img);
      bulk_commit();
      bulk_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttTmemCols);
  }
}

}  // namespace ada
