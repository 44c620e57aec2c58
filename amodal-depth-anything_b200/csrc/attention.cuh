// Fused flash-style multi-head attention for sm_100a, head_dim 64 (every DINOv2 size), no mask, no dropout.
// Replaces  softmax((q * d^-0.5) k^T) v  of attention.py:53-59 (== xformers memory_efficient_attention, attention.py:74-77)
// without materialising the N x N score matrix.
//
// One CTA = 128 queries of one (image, head); Q/K/V are read in place from the [B, N, 3, H, 64] QKV GEMM output
// through one 3-D TMA map (out-of-range tokens are zero-filled by the hardware).
//   warp 0     : TMA producer (Q once, then a 2-stage K ring and a 2-stage V ring)
//   warp 1     : tcgen05 issuer.  S = Q K^T  (128x128x64, both operands K-major)  -> TMEM cols [0,128)
//                                 O_j = P V  (128x64x128, P K-major from smem, V MN-major) -> TMEM cols [128,192)
//   warps 2..5 : softmax. Each thread owns one query row (= one TMEM lane), so row max / row sum need no shuffles.
//                P is written to shared memory as bf16 in the 128-byte-swizzled K-major layout the MMA expects;
//                the running output lives in registers: O = O * alpha + O_j.
// Two CTAs fit per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
#pragma once
#include "ptx.cuh"

namespace ada {

constexpr int kAttThreads = 192;
constexpr int kAttQ = 128, kAttKV = 128, kAttD = 64;
constexpr int kAttSmemBytes = 16384 /*Q*/ + 2 * 16384 /*K*/ + 2 * 16384 /*V*/ + 32768 /*P*/ + 256 /*barriers*/;
constexpr int kAttTmemCols = 256;

struct AttArgs {
  int B, N, heads, D;        // D = heads * 64
  __nv_bfloat16* out;        // [B*N, D]
  float scale_log2e;         // d^-0.5 * log2(e)
};

__global__ void __launch_bounds__(kAttThreads, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttArgs a) {
  extern __shared__ __align__(1024) uint8_t att_smem[];
  const uint32_t sbase = smem_u32(att_smem);
  const uint32_t sQ = sbase, sK = sbase + 16384, sV = sbase + 49152, sP = sbase + 81920;
  const uint32_t bar = sbase + 114688;
  const uint32_t q_full = bar, s_full = bar + 8, p_full = bar + 16, o_full = bar + 24;
  auto k_full = [&](int s) { return bar + 32 + 8u * s; };
  auto k_empty = [&](int s) { return bar + 48 + 8u * s; };
  auto v_full = [&](int s) { return bar + 64 + 8u * s; };
  auto v_empty = [&](int s) { return bar + 80 + 8u * s; };
  const uint32_t tmem_ptr_smem = bar + 96;

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
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
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
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      mbar_expect_tx(q_full, 16384);
      tma_load_3d(sQ, &tmap_qkv, q_full, head * kAttD, q0, img);
    }
    for (int j = 0; j < num_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(k_empty(s), ph ^ 1u, 0x500 + s);
      if (lane == 0) {
        mbar_expect_tx(k_full(s), 16384);
        tma_load_3d(sK + s * 16384, &tmap_qkv, k_full(s), a.D + head * kAttD, j * kAttKV, img);
      }
      mbar_wait(v_empty(s), ph ^ 1u, 0x510 + s);
      if (lane == 0) {
        mbar_expect_tx(v_full(s), 16384);
        tma_load_3d(sV + s * 16384, &tmap_qkv, v_full(s), 2 * a.D + head * kAttD, j * kAttKV, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B = V is MN-major (d contiguous)
    auto issue_s = [&](int j) {
      const int s = j & 1;
      mbar_wait(k_full(s), (j >> 1) & 1, 0x520 + s);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_smem_desc_sw128(sQ + k * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sK + s * 16384 + k * 32, 16, 1024);
          umma_bf16_ss(tS, da, db, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(k_empty(s));
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 0x530);
    issue_s(0);
    for (int j = 0; j < num_kv; ++j) {
      const int s = j & 1;
      mbar_wait(p_full, j & 1, 0x540);  // S(j) consumed, O_{j-1} consumed, P(j) in smem
      if (j + 1 < num_kv) issue_s(j + 1);
      mbar_wait(v_full(s), (j >> 1) & 1, 0x550 + s);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = make_smem_desc_sw128(sP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sV + s * 16384 + kk * 2048, 0, 1024);
          umma_bf16_ss(tO, da, db, idesc_o, kk > 0 ? 1u : 0u);
        }
        umma_commit(v_empty(s));
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax / output warps
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float c = a.scale_log2e;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 0.f;
    const uint32_t p_row = sP + row * 128;
    const uint32_t sw = static_cast<uint32_t>(row & 7);

    for (int j = 0; j < num_kv; ++j) {
      const int kv_valid = min(kAttKV, a.N - j * kAttKV);
      mbar_wait(s_full, j & 1, 0x560);
      tc_fence_after();
      // pass 1: row max over this KV tile
      float mx = m_run;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld32(tS + lane_off + cc * 32, r);
        tmem_ld_wait();
        if (kv_valid == kAttKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float alpha = fast_exp2((m_run - mx) * c);  // first tile: exp2(-inf) = 0
      m_run = mx;
      const float mc = mx * c;
      // fold the previous tile's P V into the running output
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1, 0x570);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld32(tO + lane_off + h * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha_pending, __uint_as_float(r[i]));
        }
      }
      alpha_pending = alpha;
      // pass 2: P = exp2(S*c - m*c) -> bf16 -> swizzled smem; row sum in fp32
      float rs = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld32(tS + lane_off + cc * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), c, -mc));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), c, -mc));
          if (cc * 32 + i >= kv_valid) p0 = 0.f;
          if (cc * 32 + i + 1 >= kv_valid) p1 = 0.f;
          rs += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        const uint32_t sub = p_row + (cc >> 1) * 16384;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t chunk = static_cast<uint32_t>((cc & 1) * 4 + i);
          const uint32_t addr = sub + ((chunk ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                       "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                       : "memory");
        }
      }
      l_run = fmaf(l_run, alpha, rs);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // last P V
    mbar_wait(o_full, (num_kv - 1) & 1, 0x580);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t r[32];
      tmem_ld32(tO + lane_off + h * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha_pending, __uint_as_float(r[i]));
    }
    const int qi = q0 + row;
    if (qi < a.N) {
      const float inv = 1.0f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(a.out + (static_cast<long long>(img) * a.N + qi) * a.D + head * kAttD);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dst[i] = make_uint4(pack_bf16x2(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                            pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv),
                            pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
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
