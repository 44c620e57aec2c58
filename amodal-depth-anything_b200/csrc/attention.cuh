// Fused flash-style multi-head attention for sm_100a, head_dim 64 (every DINOv2 size), no mask, no dropout.
// Replaces  softmax((q * d^-0.5) k^T) v  of attention.py:53-59 (== xformers memory_efficient_attention, attention.py:74-77)
// without materialising the N x N score matrix.
//
// One CTA = 256 queries (two 128-row query tiles A and B) of one (image, head); Q/K/V are read in place from the
// [B, N, 3, H, 64] QKV GEMM output through a 3-D TMA map (out-of-range tokens are zero-filled by the hardware).
// Both query tiles consume the SAME K/V tiles from shared memory: with one 128-query tile per CTA the kernel was bound by
// L2 -> SM operand traffic (every CTA streams the whole K and V of its head: 10.4 TB/s measured with all math removed).
//   warp 0      : TMA producer (Q_A, Q_B once, then 3-stage K and V rings of 128-key tiles)
//   warp 1      : tcgen05 issuer, event driven over the two tiles:
//                    S_t = Q_t K^T  (128x128x64, both operands K-major)             -> TMEM cols [128 t, 128 t + 128)
//                    O_t += P_t V   (128x64x128, P from TMEM, V MN-major from smem) -> TMEM cols [384 + 64 t, ...)
//   warps 4..7  : softmax for tile A,  warps 8..11 : softmax for tile B. Each thread owns one query row (= one TMEM lane):
//                 the 128-wide score row is pulled into registers (four tcgen05.ld, one wait) and S is released to the
//                 issuer before the exponentials start; bf16 P goes back to TMEM (cols [256 + 64 t, ...)) with
//                 tcgen05.st and is consumed by the P V MMA straight from there. Row max / row sum need no shuffles.
// The running output stays in TMEM across KV tiles. Rows are kept relative to a *stale* maximum: the accumulator is only
// rescaled (tcgen05.ld -> scale -> tcgen05.st) when some row's maximum grows by more than 2^8, which happens in the first
// tile or two; p may then reach 256, harmless in fp32/bf16, and O / l is exact either way.
// While tile A's warps exponentiate, the tensor pipe works for tile B and vice versa.
#pragma once
#include "ptx.cuh"

namespace ada {

constexpr int kAttThreads = 384;
constexpr int kAttQ = 128, kAttKV = 128, kAttD = 64;
constexpr int kAttStages = 3;
constexpr int kAttTileBytes = kAttKV * kAttD * 2;  // 16 KB
constexpr int kAttSmemBytes = 2 * 16384 /*Q_A,Q_B*/ + 2 * kAttStages * kAttTileBytes /*K,V*/ + 256 /*barriers*/;
constexpr int kAttTmemCols = 512;
constexpr float kAttRescaleLog2 = 8.0f;  // rescale O only when a row max grows by more than 2^8

struct AttArgs {
  int B, N, heads, D;        // D = heads * 64
  float scale_log2e;         // d^-0.5 * log2(e)
};

// VARIANT is a measurement knob (env ADA_ATT_VARIANT): 0 = product, 1 = 3/8 of the exponentials on the FMA pipe,
// 2 = exponentials replaced by a copy (timing skeleton only, wrong results).
template <int VARIANT>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                         const AttArgs a) {
  extern __shared__ __align__(1024) uint8_t att_smem[];
  const uint32_t sbase = smem_u32(att_smem);
  const uint32_t sQ = sbase, sK = sbase + 32768, sV = sK + kAttStages * kAttTileBytes;
  const uint32_t bar = sV + kAttStages * kAttTileBytes;
  auto q_full = [&](int t) { return bar + 8u * t; };
  auto s_full = [&](int t) { return bar + 16 + 8u * t; };
  auto s_free = [&](int t) { return bar + 32 + 8u * t; };
  auto p_full = [&](int t) { return bar + 48 + 8u * t; };
  auto pv_done = [&](int t) { return bar + 64 + 8u * t; };
  auto k_full = [&](int s) { return bar + 80 + 8u * s; };
  auto k_empty = [&](int s) { return bar + 112 + 8u * s; };
  auto v_full = [&](int s) { return bar + 144 + 8u * s; };
  auto v_empty = [&](int s) { return bar + 176 + 8u * s; };
  const uint32_t tmem_ptr_smem = bar + 208;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * kAttQ, head = blockIdx.y, img = blockIdx.z;
  const int num_kv = (a.N + kAttKV - 1) / kAttKV;
  const int n_tiles = (q0 + kAttQ < a.N) ? 2 : 1;  // tile B is skipped entirely when it lies past the sequence

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {  // the swizzled layouts below assume a 1 KB aligned window
      g_dev_error[0] = 0xA11;
      __trap();
    }
    tma_prefetch_desc(&tmap_qkv);
    for (int t = 0; t < 2; ++t) {
      mbar_init(q_full(t), 1);
      mbar_init(s_full(t), 1);
      mbar_init(s_free(t), 128);
      mbar_init(p_full(t), 128);
      mbar_init(pv_done(t), 1);
    }
    for (int s = 0; s < kAttStages; ++s) {
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
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        mbar_expect_tx(q_full(t), 16384);
        tma_load_3d(sQ + t * 16384, &tmap_qkv, q_full(t), head * kAttD, q0 + t * kAttQ, img);
      }
    }
    for (int j = 0; j < num_kv; ++j) {
      const int s = j % kAttStages;
      const uint32_t ph = (j / kAttStages) & 1;
      mbar_wait(k_empty(s), ph ^ 1u, 0x500 + s);
      if (lane == 0) {
        mbar_expect_tx(k_full(s), kAttTileBytes);
        tma_load_3d(sK + s * kAttTileBytes, &tmap_qkv, k_full(s), a.D + head * kAttD, j * kAttKV, img);
      }
      mbar_wait(v_empty(s), ph ^ 1u, 0x510 + s);
      if (lane == 0) {
        mbar_expect_tx(v_full(s), kAttTileBytes);
        tma_load_3d(sV + s * kAttTileBytes, &tmap_qkv, v_full(s), 2 * a.D + head * kAttD, j * kAttKV, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (event driven over both tiles)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B = V is MN-major (d contiguous)
    int ns[2] = {0, 0};    // next S tile index to issue, per query tile
    int npv[2] = {0, 0};   // next P V index to issue
    bool qok[2] = {false, false};
    const long long t_start = clock64();
    int remaining = n_tiles * 2 * num_kv;
    while (remaining > 0) {
      bool progress = false;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t >= n_tiles) continue;
        // ---- S_t(j) = Q_t K_j^T : needs Q_t, K_j, and S_t free (softmax has pulled S_t(j-1) into registers)
        {
          const int j = ns[t];
          if (j < num_kv) {
            const int s = j % kAttStages;
            if (!qok[t]) qok[t] = mbar_try_wait(q_full(t), 0);
            const bool ready = qok[t] && mbar_try_wait(k_full(s), (j / kAttStages) & 1) &&
                               (j == 0 || mbar_try_wait(s_free(t), (j - 1) & 1));
            if (ready) {
              tc_fence_after();
              if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = make_smem_desc_sw128(sQ + t * 16384 + k * 32, 16, 1024);
                  const uint64_t db = make_smem_desc_sw128(sK + s * kAttTileBytes + k * 32, 16, 1024);
                  umma_bf16_ss(tmem_base + t * 128, da, db, idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(s_full(t));
                // the K stage is free once BOTH tiles' S(j) have been issued (commit covers every prior MMA)
                if (n_tiles == 1 || ns[t ^ 1] > j) umma_commit(k_empty(s));
              }
              __syncwarp();
              ns[t] = j + 1;
              --remaining;
              progress = true;
            }
          }
        }
        // ---- O_t += P_t(j) V_j : needs P_t(j) in TMEM (softmax done with tile j) and V_j
        {
          const int j = npv[t];
          if (j < num_kv && j < ns[t]) {
            const int s = j % kAttStages;
            const bool ready = mbar_try_wait(p_full(t), j & 1) && mbar_try_wait(v_full(s), (j / kAttStages) & 1);
            if (ready) {
              tc_fence_after();
              if (lane == 0) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {  // 16 keys per MMA = 8 packed TMEM columns of P
                  const uint64_t db = make_smem_desc_sw128(sV + s * kAttTileBytes + kk * 2048, 0, 1024);
                  umma_bf16_ts(tmem_base + 384 + t * 64, tmem_base + 256 + t * 64 + kk * 8, db, idesc_o,
                               (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(pv_done(t));
                if (n_tiles == 1 || npv[t ^ 1] > j) umma_commit(v_empty(s));
              }
              __syncwarp();
              npv[t] = j + 1;
              --remaining;
              progress = true;
            }
          }
        }
      }
      if (!progress && clock64() - t_start > 4000000000LL) {  // protocol bug: fail loudly instead of hanging
        g_dev_error[0] = 0x5FF;
        g_dev_error[1] = blockIdx.x;
        g_dev_error[2] = static_cast<unsigned>(ns[0] | (npv[0] << 8) | (ns[1] << 16) | (npv[1] << 24));
        __trap();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / output warps (tile t = warpgroup)
    const int t = (warp - 4) >> 2;
    if (t < n_tiles) {
      const int qd = warp & 3;
      const int row = qd * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
      const uint32_t tS = tmem_base + t * 128 + lane_off, tP = tmem_base + 256 + t * 64 + lane_off,
                     tO = tmem_base + 384 + t * 64 + lane_off;
      const float c = a.scale_log2e;
      float m_used = -INFINITY, l_run = 0.f;

      for (int j = 0; j < num_kv; ++j) {
        const int kv_valid = min(kAttKV, a.N - j * kAttKV);
        mbar_wait(s_full(t), j & 1, 0x560 + t);
        tc_fence_after();
        uint32_t s0[32], s1[32], s2[32], s3[32];
        tmem_ld32(tS, s0);
        tmem_ld32(tS + 32, s1);
        tmem_ld32(tS + 64, s2);
        tmem_ld32(tS + 96, s3);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free(t));  // S_t(j) now lives in registers
        if (kv_valid < kAttKV) {  // ragged last tile: keys past N are zero-filled by TMA -> mask them out
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= kv_valid) s0[i] = 0xff800000u;
            if (32 + i >= kv_valid) s1[i] = 0xff800000u;
            if (64 + i >= kv_valid) s2[i] = 0xff800000u;
            if (96 + i >= kv_valid) s3[i] = 0xff800000u;
          }
        }
        // 16 independent max chains
        float tm[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tm[i] = fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 4]));
          tm[4 + i] = fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 4]));
          tm[8 + i] = fmaxf(__uint_as_float(s2[i]), __uint_as_float(s2[i + 4]));
          tm[12 + i] = fmaxf(__uint_as_float(s3[i]), __uint_as_float(s3[i + 4]));
        }
#pragma unroll
        for (int i = 8; i < 32; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tm[k] = fmaxf(tm[k], __uint_as_float(s0[i + k]));
            tm[4 + k] = fmaxf(tm[4 + k], __uint_as_float(s1[i + k]));
            tm[8 + k] = fmaxf(tm[8 + k], __uint_as_float(s2[i + k]));
            tm[12 + k] = fmaxf(tm[12 + k], __uint_as_float(s3[i + k]));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) tm[i] = fmaxf(tm[i], tm[i + 8]);
#pragma unroll
        for (int i = 0; i < 4; ++i) tm[i] = fmaxf(tm[i], tm[i + 4]);
        const float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));

        // ---- running (stale) maximum: decide now, in registers; the accumulator itself is rescaled further down
        float sc = 1.0f;
        bool rescale = false;
        if (j == 0) {
          m_used = tmax;
        } else {
          const bool grow = (tmax - m_used) * c > kAttRescaleLog2;
          rescale = __any_sync(0xffffffffu, grow);  // rare (first tile or two)
          if (rescale) {
            const float m_new = fmaxf(m_used, tmax);
            sc = fast_exp2((m_used - m_new) * c);
            m_used = m_new;
            l_run *= sc;
          }
        }
        const float mc = m_used * c;
        float rs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // independent row-sum chains
        auto expo = [&](const uint32_t (&sa)[32], const uint32_t (&sb)[32], uint32_t (&pk)[32]) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float x0 = fmaf(__uint_as_float(h ? sb[i] : sa[i]), c, -mc);
              const float x1 = fmaf(__uint_as_float(h ? sb[i + 1] : sa[i + 1]), c, -mc);
              const int e = (i >> 1) & 7;
              float p0, p1;
              if constexpr (VARIANT == 2) {
                p0 = x0;
                p1 = x1;
              } else if constexpr (VARIANT == 1) {
                p0 = (e == 1 || e == 4 || e == 6) ? exp2_fma(x0) : fast_exp2(x0);
                p1 = (e == 2 || e == 4 || e == 7) ? exp2_fma(x1) : fast_exp2(x1);
              } else {
                p0 = fast_exp2(x0);
                p1 = fast_exp2(x1);
              }
              (void)e;
              rs[(i >> 1) & 3] += p0;
              rs[4 + ((i >> 1) & 3)] += p1;
              pk[h * 16 + (i >> 1)] = pack_bf16x2(p0, p1);
            }
          }
        };
        uint32_t pk[32];
        expo(s0, s1, pk);  // first half of the row before touching TMEM: hides the wait for P(j-1) V(j-1) below
        if (j > 0) {       // P_t(j-1) V(j-1) must have retired before P_t is overwritten / O_t may be rescaled
          mbar_wait(pv_done(t), (j - 1) & 1, 0x570 + t);
          tc_fence_after();
          if (rescale) {  // bring this warp's 32 accumulator rows to the new maxima
#pragma unroll 1
            for (int h = 0; h < 8; ++h) {  // 8 columns at a time: rare path, keep its register footprint small
              uint32_t r[8];
              tmem_ld8(tO + h * 8, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * sc);
              tmem_st8(tO + h * 8, r);
            }
          }
        }
        tmem_st32(tP, pk);
        expo(s2, s3, pk);
        tmem_st32(tP + 32, pk);
        tmem_st_wait();
        l_run += ((rs[0] + rs[1]) + (rs[2] + rs[3])) + ((rs[4] + rs[5]) + (rs[6] + rs[7]));
        tc_fence_before();
        mbar_arrive(p_full(t));
      }
      // ---- epilogue: O / l -> bf16 -> swizzled smem (Q_t is dead once the last S_t MMA has retired) -> one TMA store
      mbar_wait(pv_done(t), (num_kv - 1) & 1, 0x580 + t);
      tc_fence_after();
      const float inv = 1.0f / l_run;
      const uint32_t o_row = sQ + t * 16384 + static_cast<uint32_t>(row) * 128u;
      const uint32_t o_sw = static_cast<uint32_t>(row & 7);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        tmem_ld32(tO + h * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t chunk = static_cast<uint32_t>(h * 4 + i);
          st_shared_v4(o_row + ((chunk ^ o_sw) << 4),
                       pack_bf16x2(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv),
                       pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv),
                       pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv),
                       pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv));
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if ((threadIdx.x & 127) == 0) {  // first thread of this warpgroup; rows past N are clipped by the tensor map
        tma_store_3d(&tmap_out, sQ + t * 16384, head * kAttD, q0 + t * kAttQ, img);
        bulk_commit();
        bulk_wait<0>();
      }
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
