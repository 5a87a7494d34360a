// attn_long.cu — key-blocked attention for ANY sequence length on tcgen05 (head_dim 64), sm_100a.
//
// attn.cu keeps the whole key range of a (sequence, head) in TMEM, which caps Lk at 256.  The fine-tuning
// configurations of X²-VLM run 384 px (N = 577 patch tokens) and 768 px (N = 2305) images through the BEiT blocks
// (models/beit2.py:135-159, dense [H,N,N] relative-position bias) and let 40-token captions attend to those 577 /
// 2305 image tokens in the fusion layers (models/xbert.py:364-410).  These kernels stream the keys in blocks of 128:
//
//   forward  (attn_long_fwd_kernel): a CTA owns TWO 128-row query tiles of one (b, h); warps 0-3 work on tile A,
//     warps 4-7 on tile B (thread == query row == TMEM lane, so row max / sum need no shuffles) and warp 8 is the
//     TMA producer that streams K/V blocks through a 2-stage mbarrier ring shared by both tiles.  Per key block a
//     warpgroup's elected thread issues S = Q·K_jᵀ (tcgen05.mma into TMEM), the group applies scale + streamed bias
//     + mask, the ONLINE softmax (running max with FlashAttention-4's lazy rescale: the accumulator is only
//     rescaled when the max grows by more than 2^8), writes P (bf16, 128B-swizzled K-major) and issues O += P·V_j.
//     While one group is in its softmax the tensor pipe runs the other group's MMAs (two query tiles in flight).
//   backward (attn_long_bwd_kernel): a CTA owns ONE 128-key block of one (b, h) and streams the query tiles
//     (Q, dO double-buffered by TMA): S and dP by two MMA chains, P / dS from the saved log-sum-exp, then
//     dV += Pᵀ·dO and dK += dSᵀ·Q stay in TMEM for the whole kernel while the tile's dQ contribution dS·K is
//     drained from TMEM and added to an fp32 workspace with vector reductions (red.global.add.v4.f32) — the only
//     cross-CTA exchange; a streaming kernel turns the workspace into the bf16 dQ.
// Same Philox dropout stream, bias / mask layout and dS export as attn.cu.
#include "attn_common.cuh"

namespace x2k {
namespace {

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int LF_THREADS = 288;  // 2 softmax / MMA-issue warpgroups + 1 producer warp
constexpr int LF_NS = 2;         // K/V ring stages
constexpr int LF_Q = 0;                          // 2 x 16 KB
constexpr int LF_K = 32768;                      // LF_NS x 16 KB
constexpr int LF_V = LF_K + LF_NS * 16384;       // LF_NS x 16 KB
constexpr int LF_P = LF_V + LF_NS * 16384;       // 2 x 32 KB
constexpr int LF_STAGE = LF_P + 65536;           // bias staging, 8 warps x 4 KB
constexpr int LF_BARS = LF_STAGE + 32768;
constexpr int LF_SMEM = LF_BARS + 256 + 1024;
constexpr float kRescaleThreshold = 8.0f;        // log2 units: rescale the accumulator only when the row max grew by > 2^8

__global__ void __launch_bounds__(LF_THREADS, 1)
attn_long_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + LF_BARS);
  uint64_t* bar_full = bar_q + 1;            // [LF_NS]
  uint64_t* bar_empty = bar_full + LF_NS;    // [LF_NS]
  uint64_t* bar_s = bar_empty + LF_NS;       // [2]
  uint64_t* bar_o = bar_s + 2;               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_o + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int tile0 = blockIdx.x * 2;
  const int n_tiles = (p.Lq + 127) >> 7;
  const int n_act = min(2, n_tiles - tile0);  // query tiles (= active warpgroups) of this CTA
  const int nkb = (p.Lk + 127) >> 7;
  const int kvb = p.kv_index ? p.kv_index[b] : b;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < LF_NS; ++i) {
      mbar_init(bar_full + i, 1);
      mbar_init(bar_empty + i, n_act);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_s + i, 1);
      mbar_init(bar_o + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, n_act * 16384);
      for (int t = 0; t < n_act; ++t)
        tma_load_2d(smem + LF_Q + t * 16384, &tmap_q, bar_q, h * 64, b * p.Lq + (tile0 + t) * 128);
      for (int j = 0; j < nkb; ++j) {
        const int s = j % LF_NS;
        if (j >= LF_NS) mbar_wait(bar_empty + s, ((j / LF_NS) - 1) & 1);
        mbar_arrive_expect_tx(bar_full + s, 2 * 16384);
        tma_load_2d(smem + LF_K + s * 16384, &tmap_k, bar_full + s, h * 64, kvb * p.Lk + j * 128);
        tma_load_2d(smem + LF_V + s * 16384, &tmap_v, bar_full + s, h * 64, kvb * p.Lk + j * 128);
      }
    }
    __syncwarp();
  } else if ((warp >> 2) < n_act) {
    // ===== one warpgroup per query tile =====
    const int wg = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const int tile = tile0 + wg;
    const bool elected = (threadIdx.x & 127) == 0;
    const uint32_t tS = tmem + wg * 192, tO = tS + 128;
    const uint32_t trowS = tS + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t trowO = tO + (static_cast<uint32_t>(quad * 32) << 16);
    const int q = tile * 128 + row;
    const bool qvalid = q < p.Lq;
    const int q_warp0 = tile * 128 + quad * 32;
    const bool warp_live = q_warp0 < p.Lq;  // warp-uniform
    const float* bias_blk = (p.bias && warp_live) ? p.bias + h * p.bias_h_stride + static_cast<int64_t>(q_warp0) * p.bias_q_stride : nullptr;
    const float* mask_row = (p.mask && qvalid) ? p.mask + b * p.mask_b_stride + q * p.mask_q_stride : nullptr;
    const uint32_t stage_addr = sbase + LF_STAGE + warp * 4096;
    const uint32_t sP = sbase + LF_P + wg * 32768;
    const uint32_t aq = sbase + LF_Q + wg * 16384;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
    const DropCfg dc = make_drop(p.dropout_p);
    const uint64_t doff = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
    const uint64_t drop_base = (static_cast<uint64_t>(b * p.H + h) * p.Lq + q) * p.Lk_pad;

    auto issue_s = [&](int j) {  // elected thread: S = Q · K_jᵀ once block j has landed
      const int s = j % LF_NS;
      mbar_wait(bar_full + s, (j / LF_NS) & 1);
      tc_fence_after();
      const uint32_t ak = sbase + LF_K + s * 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tS, make_smem_desc(aq + k * 32, 16, 1024), make_smem_desc(ak + k * 32, 16, 1024), idesc_s, k != 0);
      umma_commit(bar_s + wg);
    };
    if (elected) {
      mbar_wait(bar_q, 0);
      issue_s(0);
    }
    __syncwarp();

    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkb; ++j) {
      const int s = j % LF_NS;
      const int kbase = j * 128;
      const int nkc = (min(p.Lk - kbase, 128) + 15) >> 4;  // live 16-column chunks of this block
      const int nunit = (nkc + 1) >> 1;                    // 32-column units
      // S_j complete => every earlier MMA of this group (incl. O += P·V of block j-1) is complete as well
      mbar_wait_warp(bar_s + wg, j & 1);
      tc_fence_after();

      // ---- pass 1: t = scale·qk + bias + mask (log2 domain) back into TMEM, block row max ----
      float mx = -INFINITY;
      if (warp_live) {
        for (int u = 0; u < nunit; ++u) {
          const bool two = (2 * u + 1) < nkc;
          uint32_t s0[16], s1[16];
          tmem_ld_32x16(trowS + u * 32, s0);
          if (two) tmem_ld_32x16(trowS + u * 32 + 16, s1);
          float add[32];
          additive32(p, bias_blk, p.Lq - q_warp0, mask_row, kbase + u * 32, stage_addr, lane, add);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float t = (kbase + u * 32 + i < p.Lk) ? fmaf(__uint_as_float(s0[i]), p.scale_log2, add[i]) : -INFINITY;
            mx = fmaxf(mx, t);
            s0[i] = __float_as_uint(t);
          }
          tmem_st_32x16(trowS + u * 32, s0);
          if (two) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = (kbase + u * 32 + 16 + i < p.Lk) ? fmaf(__uint_as_float(s1[i]), p.scale_log2, add[16 + i]) : -INFINITY;
              mx = fmaxf(mx, t);
              s1[i] = __float_as_uint(t);
            }
            tmem_st_32x16(trowS + u * 32 + 16, s1);
          }
        }
        tmem_wait_st();
      }
      // ---- online softmax: lazy accumulator rescale ----
      float alpha = 1.0f;
      const float m_new = fmaxf(m_ref, mx);
      if (j == 0) {
        m_ref = m_new;
      } else if (m_new > m_ref + kRescaleThreshold) {
        alpha = fast_exp2(m_ref - m_new);  // m_ref == -inf -> 0: nothing has been accumulated for this row yet
        m_ref = m_new;
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[16];
          tmem_ld_32x16(trowO + c * 16, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x16(trowO + c * 16, o);
        }
        tmem_wait_st();
      }
      l_run *= alpha;
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;

      // ---- pass 2: P = exp2(t - m) -> bf16, K-major 128B-swizzled smem tile ----
      if (warp_live) {
        float sum = 0.f;
        for (int c0 = 0; c0 < nkc; c0 += 2) {
          uint32_t sv[2][16];
          tmem_ld_32x16(trowS + c0 * 16, sv[0]);
          if (c0 + 1 < nkc) tmem_ld_32x16(trowS + (c0 + 1) * 16, sv[1]);
          tmem_wait_ld();
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            if (c0 + v < nkc) {
              const int c = c0 + v;
              float pr[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                pr[i] = fast_exp2(__uint_as_float(sv[v][i]) - m_use);
                sum += pr[i];
              }
              if (p.dropout_p > 0.f) {
#pragma unroll
                for (int i0 = 0; i0 < 16; i0 += 8) {
                  float k[8];
                  drop8(p.seed, doff, (drop_base + kbase + c * 16 + i0) >> 3, dc, k);
#pragma unroll
                  for (int i = 0; i < 8; ++i) pr[i0 + i] *= k[i];
                }
              }
              st_shared_v4(sP + swz_off(row, c * 16), pack_bf16x2(pr[0], pr[1]), pack_bf16x2(pr[2], pr[3]),
                           pack_bf16x2(pr[4], pr[5]), pack_bf16x2(pr[6], pr[7]));
              st_shared_v4(sP + swz_off(row, c * 16 + 8), pack_bf16x2(pr[8], pr[9]), pack_bf16x2(pr[10], pr[11]),
                           pack_bf16x2(pr[12], pr[13]), pack_bf16x2(pr[14], pr[15]));
            }
          }
        }
        l_run += sum;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1 + wg, 128);  // the whole group has read S_j and written P_j
      if (elected) {
        tc_fence_after();
        const uint32_t av = sbase + LF_V + s * 16384;
        for (int ks = 0; ks < nkc; ++ks)
          umma_bf16(tO, make_smem_desc(sP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                    make_smem_desc(av + ks * 2048, 8192, 1024), idesc_o, (j | ks) != 0);
        umma_commit(bar_empty + s);  // K_j / V_j are free once S_j and this P·V have run
        if (j + 1 < nkb) issue_s(j + 1);
        else umma_commit(bar_o + wg);
      }
      __syncwarp();
    }
    // ---- O / l -> bf16, log-sum-exp ----
    mbar_wait_warp(bar_o + wg, 0);
    tc_fence_after();
    if (warp_live) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* dst = p.o + (static_cast<int64_t>(b) * p.Lq + q) * p.ld_o + h * 64;
      store_row_bf16<4>(dst, trowO, inv, qvalid);
      if (qvalid) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Lq + q] = ((m_ref == -INFINITY) ? 0.f : m_ref) + log2f(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// backward: grid (key blocks, H, B), 256 threads, 1 CTA / SM
// ---------------------------------------------------------------------------------------------
constexpr int LB_Q = 0, LB_DO = 32768;            // 2 stages x 16 KB each
constexpr int LB_K = 65536, LB_V = 81920, LB_P = 98304, LB_DS = 131072;
constexpr int LB_STAGE = 163840;                  // bias staging, 8 warps x 4 KB
constexpr int LB_BARS = LB_STAGE + 32768;
constexpr int LB_SMEM = LB_BARS + 256 + 1024;
constexpr uint32_t LTM_S = 0, LTM_DP = 128, LTM_DQ = 256, LTM_DK = 320, LTM_DV = 384;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_long_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     const AttnParams p, float* __restrict__ dq_ws) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + LB_BARS);  // [2]
  uint64_t* bar_kv = bar_q + 2;
  uint64_t* bar_mma = bar_q + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_q + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvb = p.kv_index ? p.kv_index[b] : b;
  const int nqb = (p.Lq + 127) >> 7;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_q + 1, 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t stage_addr = sbase + LB_STAGE + warp * 4096;

  auto load_q = [&](int qb) {  // thread 0: Q / dO tile qb -> stage qb & 1
    const int st = qb & 1;
    mbar_arrive_expect_tx(bar_q + st, 2 * 16384);
    tma_load_2d(smem + LB_Q + st * 16384, &tmap_q, bar_q + st, h * 64, b * p.Lq + qb * 128);
    tma_load_2d(smem + LB_DO + st * 16384, &tmap_do, bar_q + st, h * 64, b * p.Lq + qb * 128);
  };
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_kv, 2 * 16384);
    tma_load_2d(smem + LB_K, &tmap_k, bar_kv, h * 64, kvb * p.Lk + kb * 128);
    tma_load_2d(smem + LB_V, &tmap_v, bar_kv, h * 64, kvb * p.Lk + kb * 128);
    load_q(0);
  }

  const DropCfg dc = make_drop(p.dropout_p);
  const uint64_t doff = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
  uint32_t mma_phase = 0;
  const uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);
  const uint32_t idesc_dkv = make_idesc_bf16(128, 64, 1, 1);
  // valid keys of this block in 16-column chunks: S / dP are only formed (N = nkc*16) and consumed up to there
  const int nkc = (min(p.Lk - kb * 128, 128) + 15) >> 4;
  const int nunit = (nkc + 1) >> 1;  // 32-column units, split between the two halves
  const int u_begin = half == 0 ? 0 : (nunit + 1) >> 1;
  const int u_end = half == 0 ? (nunit + 1) >> 1 : nunit;
  const uint32_t idesc_skb = make_idesc_bf16(128, nkc * 16, 0, 0);
  const int64_t ld_ws = static_cast<int64_t>(p.H) * 64;

  // this tile's dQ contribution (dS · K_kb, already in TMEM) -> fp32 workspace, reduced across the key-block CTAs
  auto drain_dq = [&](int qb) {
    const int q = qb * 128 + row;
    const bool live = (qb * 128 + quad * 32) < p.Lq;  // warp-uniform
    if (live) {
      uint32_t a0[16], a1[16];
      tmem_ld_32x16(trow + LTM_DQ + half * 32, a0);
      tmem_ld_32x16(trow + LTM_DQ + half * 32 + 16, a1);
      tmem_wait_ld();
      if (q < p.Lq) {
        float* dst = dq_ws + (static_cast<int64_t>(b) * p.Lq + q) * ld_ws + h * 64 + half * 32;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(dst + i, __uint_as_float(a0[i]), __uint_as_float(a0[i + 1]), __uint_as_float(a0[i + 2]), __uint_as_float(a0[i + 3]));
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(dst + 16 + i, __uint_as_float(a1[i]), __uint_as_float(a1[i + 1]), __uint_as_float(a1[i + 2]), __uint_as_float(a1[i + 3]));
      }
    }
  };

  for (int qb = 0; qb < nqb; ++qb) {
    const int st = qb & 1;
    if (threadIdx.x == 0) {
      if (qb == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q + st, (qb >> 1) & 1);
      tc_fence_after();
      const uint32_t aq = sbase + LB_Q + st * 16384, ado = sbase + LB_DO + st * 16384;
      const uint32_t ak = sbase + LB_K, av = sbase + LB_V;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + LTM_S, make_smem_desc(aq + k * 32, 16, 1024), make_smem_desc(ak + k * 32, 16, 1024), idesc_skb, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + LTM_DP, make_smem_desc(ado + k * 32, 16, 1024), make_smem_desc(av + k * 32, 16, 1024), idesc_skb, k != 0);
      umma_commit(bar_mma);
    }
    __syncwarp();
    // per-row statistics of this query tile (issued before the wait: the loads fly while the MMAs run)
    const int q = qb * 128 + row;
    const bool qvalid = q < p.Lq;
    float my_lse = 0.f, my_delta = 0.f;
    if (qvalid) {
      my_lse = __ldg(p.lse + (static_cast<int64_t>(b) * p.H + h) * p.Lq + q);
      my_delta = __ldg(p.delta + (static_cast<int64_t>(b) * p.H + h) * p.Lq + q);
    }
    mbar_wait_warp(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // S / dP of tile qb complete => the dK / dV / dQ MMAs of tile qb-1 are complete: its Q / dO stage is free
    if (threadIdx.x == 0 && qb + 1 < nqb) load_q(qb + 1);
    __syncwarp();
    if (qb > 0) drain_dq(qb - 1);

    // ---- P and dS for this (q tile, key block); thread == query row, half == column range ----
    const int q_warp0 = qb * 128 + quad * 32;
    const bool warp_live = q_warp0 < p.Lq;
    const float* bias_blk = (p.bias && warp_live) ? p.bias + h * p.bias_h_stride + static_cast<int64_t>(q_warp0) * p.bias_q_stride : nullptr;
    const float* mask_row = (p.mask && qvalid) ? p.mask + b * p.mask_b_stride + q * p.mask_q_stride : nullptr;
    const uint64_t drop_base = (static_cast<uint64_t>(b * p.H + h) * p.Lq + q) * p.Lk_pad;
    __nv_bfloat16* ds_row = (p.ds_out && qvalid) ? p.ds_out + b * p.ds_b_stride + h * p.ds_h_stride + q * p.ds_q_stride : nullptr;
    if (warp_live) {
#pragma unroll 1
      for (int u = u_begin; u < u_end; ++u) {
        const bool two = (2 * u + 1) < nkc;
        uint32_t s[2][16], dp[2][16];
        tmem_ld_32x16(trow + LTM_S + u * 32, s[0]);
        tmem_ld_32x16(trow + LTM_DP + u * 32, dp[0]);
        if (two) {
          tmem_ld_32x16(trow + LTM_S + u * 32 + 16, s[1]);
          tmem_ld_32x16(trow + LTM_DP + u * 32 + 16, dp[1]);
        }
        float add[32];
        additive32(p, bias_blk, p.Lq - q_warp0, mask_row, kb * 128 + u * 32, stage_addr, lane, add);
        tmem_wait_ld();
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (v == 1 && !two) continue;  // warp-uniform
          const int c = 2 * u + v;
          const int k0 = kb * 128 + c * 16;
          float pr[16], ds[16];
          if (qvalid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float t = fmaf(__uint_as_float(s[v][j]), p.scale_log2, add[16 * v + j]);
              pr[j] = (k0 + j < p.Lk) ? fast_exp2(t - my_lse) : 0.f;
              ds[j] = __uint_as_float(dp[v][j]);
            }
            if (p.dropout_p > 0.f) {
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                float k[8];
                drop8(p.seed, doff, (drop_base + k0 + j) >> 3, dc, k);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float pu = pr[j + i];
                  ds[j + i] = pu * (ds[j + i] * k[i] - my_delta);
                  pr[j + i] = pu * k[i];
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) ds[j] = pr[j] * (ds[j] - my_delta);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { pr[j] = 0.f; ds[j] = 0.f; }
          }
          uint32_t pk[8], dk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pk[j] = pack_bf16x2(pr[2 * j], pr[2 * j + 1]);
            dk[j] = pack_bf16x2(ds[2 * j], ds[2 * j + 1]);
          }
          const uint32_t o0 = swz_off(row, c * 16), o1 = swz_off(row, c * 16 + 8);
          st_shared_v4(sbase + LB_P + o0, pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(sbase + LB_P + o1, pk[4], pk[5], pk[6], pk[7]);
          st_shared_v4(sbase + LB_DS + o0, dk[0], dk[1], dk[2], dk[3]);
          st_shared_v4(sbase + LB_DS + o1, dk[4], dk[5], dk[6], dk[7]);
          if (ds_row) {
            *reinterpret_cast<uint4*>(ds_row + k0) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
            *reinterpret_cast<uint4*>(ds_row + k0 + 8) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
          }
        }
      }
    } else {
      // rows beyond Lq feed dK / dV through the MN-major operands: they must be exact zeros, not stale data
      for (int c = half; c < nkc; c += 2) {
        const uint32_t o0 = swz_off(row, c * 16), o1 = swz_off(row, c * 16 + 8);
        st_shared_v4(sbase + LB_P + o0, 0u, 0u, 0u, 0u);
        st_shared_v4(sbase + LB_P + o1, 0u, 0u, 0u, 0u);
        st_shared_v4(sbase + LB_DS + o0, 0u, 0u, 0u, 0u);
        st_shared_v4(sbase + LB_DS + o1, 0u, 0u, 0u, 0u);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (threadIdx.x == 0) {
      const uint32_t aq = sbase + LB_Q + st * 16384, ado = sbase + LB_DO + st * 16384;
      const uint32_t ak = sbase + LB_K;
      const uint32_t ap = sbase + LB_P, ads = sbase + LB_DS;
      const int nqc = (min(p.Lq - qb * 128, 128) + 15) >> 4;  // valid query rows in 16-row groups
      // dQ_tile = dS · K_kb        (A: dS K-major over keys; B: K tile MN-major, N = 64 dims)
      for (int ks = 0; ks < nkc; ++ks)
        umma_bf16(tmem + LTM_DQ, make_smem_desc(ads + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                  make_smem_desc(ak + ks * 2048, 8192, 1024), idesc_dq, ks != 0);
      // dK += dSᵀ · Q_tile         (A: dS MN-major over keys, K = query rows; B: Q tile MN-major)
      for (int ks = 0; ks < nqc; ++ks)
        umma_bf16(tmem + LTM_DK, make_smem_desc(ads + ks * 2048, 16384, 1024), make_smem_desc(aq + ks * 2048, 8192, 1024),
                  idesc_dkv, (qb | ks) != 0);
      // dV += Pᵀ · dO_tile
      for (int ks = 0; ks < nqc; ++ks)
        umma_bf16(tmem + LTM_DV, make_smem_desc(ap + ks * 2048, 16384, 1024), make_smem_desc(ado + ks * 2048, 8192, 1024),
                  idesc_dkv, (qb | ks) != 0);
      if (qb == nqb - 1) umma_commit(bar_mma);
    }
    __syncwarp();
  }
  // ---- drain: the last tile's dQ contribution, then dK / dV of this key block (thread == key row) ----
  mbar_wait_warp(bar_mma, mma_phase);
  tc_fence_after();
  drain_dq(nqb - 1);
  {
    const int key = kb * 128 + row;
    const bool kvalid = key < p.Lk;
    const int64_t r = static_cast<int64_t>(b) * p.Lk + key;
    store_row_bf16<2>(p.dk + r * p.ld_dk + h * 64 + half * 32, trow + LTM_DK + half * 32, p.scale, kvalid);
    store_row_bf16<2>(p.dv + r * p.ld_dv + h * 64 + half * 32, trow + LTM_DV + half * 32, 1.0f, kvalid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// dq[r, c] = bf16(scale * ws[r, c]) for the H*64 head columns of every query row (8 elements per thread)
__global__ void __launch_bounds__(256)
dq_finalize_kernel(const float* __restrict__ ws, int64_t rows, int cols, float scale, __nv_bfloat16* __restrict__ dq, int64_t ld_dq) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int cpr = cols >> 3;
  if (i >= rows * cpr) return;
  const int64_t r = i / cpr;
  const int c = static_cast<int>(i - r * cpr) << 3;
  const float4 a = __ldg(reinterpret_cast<const float4*>(ws + r * cols + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(ws + r * cols + c + 4));
  uint4 v;
  v.x = pack_bf16x2(a.x * scale, a.y * scale);
  v.y = pack_bf16x2(a.z * scale, a.w * scale);
  v.z = pack_bf16x2(b.x * scale, b.y * scale);
  v.w = pack_bf16x2(b.z * scale, b.w * scale);
  *reinterpret_cast<uint4*>(dq + r * ld_dq + c) = v;
}

}  // namespace

int64_t attn_long_bwd_ws_bytes(const X2kAttnArgs& a) {
  return static_cast<int64_t>(a.B) * a.Lq * a.H * 64 * static_cast<int64_t>(sizeof(float));
}

int attn_long_fwd(const X2kAttnArgs& a, cudaStream_t stream) {
  AttnParams p;
  fill_params(a, p);
  const int n_kv = a.n_kv > 0 ? a.n_kv : a.B;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_q, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, a.k, static_cast<uint64_t>(n_kv) * a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_k, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, a.v, static_cast<uint64_t>(n_kv) * a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_v, 128, 64))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_long_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM));
    attr_set = true;
  }
  const int n_tiles = (a.Lq + 127) / 128;
  dim3 grid((n_tiles + 1) / 2, a.H, a.B);
  attn_long_fwd_kernel<<<grid, LF_THREADS, LF_SMEM, stream>>>(tq, tk, tv, p);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

int attn_long_bwd(const X2kAttnArgs& a, cudaStream_t stream) {
  X2K_REQUIRE(a.delta_ws != nullptr, "x2k_attn_bwd: the key-blocked kernel needs delta_ws (fp32 [B,H,Lq])");
  X2K_REQUIRE(a.dq_ws != nullptr, "x2k_attn_bwd: Lq or Lk > 256 needs dq_ws (x2k_attn_bwd_workspace_bytes)");
  X2K_REQUIRE((reinterpret_cast<uintptr_t>(a.dq_ws) & 15) == 0, "x2k_attn_bwd: dq_ws must be 16-byte aligned");
  X2K_REQUIRE(!a.ds_out || (a.ds_q_stride % 8 == 0 && a.ds_h_stride % 8 == 0 && a.ds_b_stride % 8 == 0 &&
                            a.ds_q_stride >= ((a.Lk + 15) & ~15)),
              "x2k_attn_bwd: ds_out strides must be multiples of 8 and cover Lk_pad");
  AttnParams p;
  fill_params(a, p);
  const int n_kv = a.n_kv > 0 ? a.n_kv : a.B;
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_q, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, a.k, static_cast<uint64_t>(n_kv) * a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_k, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, a.v, static_cast<uint64_t>(n_kv) * a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_v, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, a.d_o, static_cast<uint64_t>(a.B) * a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_do, 128, 64))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_long_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM));
    attr_set = true;
  }
  X2K_CHECK_CUDA(cudaMemsetAsync(a.dq_ws, 0, static_cast<size_t>(attn_long_bwd_ws_bytes(a)), stream));
  dim3 grid((a.Lk + 127) / 128, a.H, a.B);
  attn_long_bwd_kernel<<<grid, ATT_THREADS, LB_SMEM, stream>>>(tq, tk, tv, tdo, p, a.dq_ws);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  const int64_t rows = static_cast<int64_t>(a.B) * a.Lq;
  const int cols = a.H * 64;
  const int64_t n = rows * (cols / 8);
  dq_finalize_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(a.dq_ws, rows, cols, a.scale,
                                                                               static_cast<__nv_bfloat16*>(a.dq), a.ld_dq);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

}  // namespace x2k
