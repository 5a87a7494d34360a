// attn.cu — attention forward / backward on tcgen05 for sm_100a (head_dim 64; Lq, Lk <= 256 here, longer
// sequences are dispatched to the key-blocked kernels of attn_long.cu).
//
// One design serves BEiT self-attention (N = 197, dense per-head relative-position bias), BERT text
// self-attention (L <= 64, key / 3-D masks, probability dropout) and the fusion layers'
// cross-attention (text queries over 197 image keys, shared K/V through kv_index):
//   * Q, K, V tiles are TMA-loaded (128B swizzle) straight out of the packed projection output;
//   * S = Q·Kᵀ runs as one tcgen05.mma chain (M = 128 query rows, N = padded key count) into TMEM;
//   * softmax: thread i of the CTA owns query row i (tcgen05.ld 32x32b puts TMEM lane i in thread
//     i), so row max / sum need no shuffles; scores never leave the SM;
//   * P is written to shared memory in the K-major 128B-swizzled UMMA layout and O = P·V runs as a
//     second MMA chain with V consumed MN-major (no transpose);
//   * backward recomputes P from the saved log-sum-exp, and forms dQ, dK, dV with five MMA chains
//     per (query block, key block) tile, reusing the P / dS shared-memory tiles both K-major
//     (dQ = dS·K) and MN-major (dK = dSᵀ·Q, dV = Pᵀ·dO); TMEM holds S, dP, dQ[2], dK, dV = 512 cols.
// Replaces models/beit2.py:135-159 and models/xbert.py:364-410 (+ autograd).
#include <cstdlib>

#include "attn_common.cuh"

namespace x2k {
namespace {

// Bias block [32 rows x 32 cols] fp32 of one warp in two steps, so the L2 latency of the loads hides behind an MMA wait
// or the previous unit's math: issue (8 coalesced LDG.128 per lane, 4 rows x 128 B per instruction) ... stage (into the
// warp's 4 KB XOR-swizzled tile) ... read (every thread picks up 16 values of ITS row per 16-column chunk).
__device__ __forceinline__ void bias_issue(const float* __restrict__ base, int64_t row_stride, int rows_left, int lane,
                                           float4 (&v)[8]) {
  const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    const int rc = r < rows_left ? r : (rows_left - 1);
    v[i] = __ldg(reinterpret_cast<const float4*>(base + rc * row_stride) + ch);
  }
}
__device__ __forceinline__ void bias_stage(const float4 (&v)[8], uint32_t stage_addr, int lane) {
  const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr + r * 128 + ((ch ^ (r & 7)) << 4)), "f"(v[i].x),
                 "f"(v[i].y), "f"(v[i].z), "f"(v[i].w)
                 : "memory");
  }
  __syncwarp();
}
// the 16 values of this thread's row for 16-column chunk `v` (0 / 1) of the staged 32-column block
__device__ __forceinline__ void bias_read16(uint32_t stage_addr, int lane, int v, float (&out)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float4 w;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                 : "r"(stage_addr + lane * 128 + (((4 * v + c) ^ (lane & 7)) << 4))
                 : "memory");
    out[4 * c] = w.x; out[4 * c + 1] = w.y; out[4 * c + 2] = w.z; out[4 * c + 3] = w.w;
  }
}

// Optional phase trace (compile with -DX2K_ATTN_TRACE): clock64 stamps of one CTA, read back by x2k_debug_attn_trace.
#ifdef X2K_ATTN_TRACE
__device__ long long g_attn_trace[3][64];
__device__ long long g_attn_fwd_trace[2][64];
#define AT_TRACE(i)                                                                           \
  do {                                                                                        \
    if (blockIdx.x == 5 && blockIdx.y == 40 && (threadIdx.x == 0 || threadIdx.x == 200 || threadIdx.x == 256)) \
      g_attn_trace[threadIdx.x == 0 ? 0 : threadIdx.x == 200 ? 1 : 2][(i)] = clock64();       \
  } while (0)
#define FT_TRACE(i)                                                                           \
  do {                                                                                        \
    if (blockIdx.x == 0 && blockIdx.y == 5 && blockIdx.z == 40 && (threadIdx.x == 0 || threadIdx.x == 200)) \
      g_attn_fwd_trace[threadIdx.x == 0 ? 0 : 1][(i)] = clock64();                            \
  } while (0)
#else
#define AT_TRACE(i) do {} while (0)
#define FT_TRACE(i) do {} while (0)
#endif
// ---------------------------------------------------------------------------------------------
// forward: grid (q tiles, H, B), 256 threads, 2 CTAs / SM (TMEM 256 columns, ~101 KB smem each)
// ---------------------------------------------------------------------------------------------
constexpr int FWD_REGION0 = 65536;  // Q (16 KB) + K (<= 32 KB) during S; P (4 x 16 KB) afterwards
constexpr int FWD_REGIONV = 32768;  // bias staging (8 warps x 4 KB) during pass 1, then the V tile
constexpr int FWD_SMEM = FWD_REGION0 + FWD_REGIONV + 2048 + 1024;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sP = smem;  // aliases Q/K once S has been computed
  uint8_t* sV = smem + FWD_REGION0;
  float* s_red = reinterpret_cast<float*>(smem + FWD_REGION0 + FWD_REGIONV);  // [2 halves][128 rows] max, then sum
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + FWD_REGION0 + FWD_REGIONV + 1024);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_mma = bar_qk + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_qk + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  FT_TRACE(0);
  const int kvb = p.kv_index ? p.kv_index[b] : b;
  const int Lk_pad = p.Lk_pad;

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_qk, 16384 + Lk_pad * 128);
    tma_load_2d(sQ, &tmap_q, bar_qk, h * 64, b * p.Lq + qt * 128);
    tma_load_2d(sK, &tmap_k, bar_qk, h * 64, kvb * p.Lk);
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, Lk_pad, 0, 0);
    const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem, make_smem_desc(aq + k * 32, 16, 1024), make_smem_desc(ak + k * 32, 16, 1024), idesc, k != 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  FT_TRACE(1);
  // ---- softmax: pass 1 forms t = scale·qk + bias + mask (log2 domain), row max, writes t back to TMEM ----
  const int q = qt * 128 + row;
  const bool qvalid = q < p.Lq;
  const bool warp_live = (qt * 128 + quad * 32) < p.Lq;  // warp-uniform
  const int q_warp0 = qt * 128 + quad * 32;              // first row of this warp
  const float* bias_blk = (p.bias && warp_live) ? p.bias + h * p.bias_h_stride + static_cast<int64_t>(q_warp0) * p.bias_q_stride : nullptr;
  const float* mask_row = (p.mask && qvalid) ? p.mask + b * p.mask_b_stride + q * p.mask_q_stride : nullptr;
  const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const uint32_t stage_addr = smem_u32(sV) + warp * 4096;
  const int nchunk = Lk_pad >> 4;           // 16-column chunks
  const int nunit = (nchunk + 1) >> 1;      // 32-column units
  const int u_begin = half == 0 ? 0 : (nunit + 1) >> 1;
  const int u_end = half == 0 ? (nunit + 1) >> 1 : nunit;
  const uint32_t sP_addr = smem_u32(sP);
  float4 bv[8];  // first bias block of this warp's rows: in flight while S is being formed
  if (bias_blk && u_begin < u_end) bias_issue(bias_blk + u_begin * 32, p.bias_q_stride, p.Lq - q_warp0, lane, bv);
  mbar_wait_warp(bar_mma, 0);
  tc_fence_after();
  FT_TRACE(2);

  float mx = -INFINITY, sum = 0.f;
  if (warp_live) {
    for (int u = u_begin; u < u_end; ++u) {
      if (bias_blk) {  // this unit's bias block was requested before the wait for S / during the previous unit
        bias_stage(bv, stage_addr, lane);
        if (u + 1 < u_end) bias_issue(bias_blk + (u + 1) * 32, p.bias_q_stride, p.Lq - q_warp0, lane, bv);
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int c = 2 * u + v;
        if (c >= nchunk) break;  // warp-uniform: the last unit may hold a single chunk
        uint32_t s0[16];
        tmem_ld_32x16(trow + c * 16, s0);
        float add[16];
        if (bias_blk) {
          bias_read16(stage_addr, lane, v, add);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) add[j] = 0.f;
        }
        if (mask_row) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask_row + c * 16 + j));
            add[j] += m.x; add[j + 1] += m.y; add[j + 2] += m.z; add[j + 3] += m.w;
          }
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = (c * 16 + j < p.Lk) ? fmaf(__uint_as_float(s0[j]), p.scale_log2, add[j] * kLog2e) : -INFINITY;
          mx = fmaxf(mx, t);
          s0[j] = __float_as_uint(t);
        }
        tmem_st_32x16(trow + c * 16, s0);
      }
      if (bias_blk) __syncwarp();  // the staging tile is rewritten by the next unit
    }
    tmem_wait_st();
    s_red[half * 128 + row] = mx;
  }
  FT_TRACE(3);
  __syncthreads();  // max exchange; the bias staging area is free from here on
  FT_TRACE(4);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_v, Lk_pad * 128);
    tma_load_2d(sV, &tmap_v, bar_v, h * 64, kvb * p.Lk);  // lands while pass 2 runs
  }
  if (warp_live) {
    mx = fmaxf(mx, s_red[(half ^ 1) * 128 + row]);
    if (mx == -INFINITY) mx = 0.f;
    const DropCfg dc = make_drop(p.dropout_p);
    const uint64_t doff = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
    const uint64_t drop_base = (static_cast<uint64_t>(b * p.H + h) * p.Lq + q) * Lk_pad;
    const int c_begin = 2 * u_begin, c_end = min(nchunk, 2 * u_end);
    for (int c0 = c_begin; c0 < c_end; c0 += 2) {
      uint32_t s[2][16];
      tmem_ld_32x16(trow + c0 * 16, s[0]);
      if (c0 + 1 < c_end) tmem_ld_32x16(trow + (c0 + 1) * 16, s[1]);
      tmem_wait_ld();
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        if (c0 + v < c_end) {
          const int c = c0 + v;
          float pr[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            pr[j] = fast_exp2(__uint_as_float(s[v][j]) - mx);
            sum += pr[j];
          }
          if (p.dropout_p > 0.f) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              float k[8];
              drop8(p.seed, doff, (drop_base + c * 16 + j) >> 3, dc, k);
#pragma unroll
              for (int i = 0; i < 8; ++i) pr[j + i] *= k[i];
            }
          }
          st_shared_v4(sP_addr + swz_off(row, c * 16), pack_bf16x2(pr[0], pr[1]), pack_bf16x2(pr[2], pr[3]),
                       pack_bf16x2(pr[4], pr[5]), pack_bf16x2(pr[6], pr[7]));
          st_shared_v4(sP_addr + swz_off(row, c * 16 + 8), pack_bf16x2(pr[8], pr[9]), pack_bf16x2(pr[10], pr[11]),
                       pack_bf16x2(pr[12], pr[13]), pack_bf16x2(pr[14], pr[15]));
        }
      }
    }
  }
  FT_TRACE(5);
  __syncthreads();  // every thread has read the partner's max before the slots are reused for the sums
  if (warp_live) s_red[half * 128 + row] = sum;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- O = P · V (accumulator aliases the first 64 S columns, all S reads are complete) ----
  if (threadIdx.x == 0) {
    mbar_wait(bar_v, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t av = smem_u32(sV);
    for (int ks = 0; ks < nchunk; ++ks) {
      const uint64_t da = make_smem_desc(sP_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
      const uint64_t db = make_smem_desc(av + ks * 2048, 8192, 1024);
      umma_bf16(tmem, da, db, idesc, ks != 0);
    }
    umma_commit(bar_mma);
  }
  __syncwarp();
  FT_TRACE(6);
  mbar_wait_warp(bar_mma, 1);
  tc_fence_after();
  FT_TRACE(7);
  if (warp_live) {  // each half writes 32 of the 64 output dims of its rows
    const float tot = sum + s_red[(half ^ 1) * 128 + row];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;
    __nv_bfloat16* dst = p.o + (static_cast<int64_t>(b) * p.Lq + q) * p.ld_o + h * 64 + half * 32;
    store_row_bf16<2>(dst, trow + half * 32, inv, qvalid);
    if (qvalid && half == 0) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Lq + q] = mx + log2f(tot);
  }
  FT_TRACE(8);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
  FT_TRACE(9);
}

// ---------------------------------------------------------------------------------------------
// backward: grid (H, B), 384 threads (8 compute warps + MMA warp + store warp), 1 CTA / SM (TMEM 512 columns)
// ---------------------------------------------------------------------------------------------
constexpr int BWD_SQ = 0, BWD_SDO = 32768, BWD_SK = 65536, BWD_SV = 98304, BWD_SP = 131072, BWD_SDS = 163840;
constexpr int BWD_STAGE = 196608;  // bias staging, 8 warps x 4 KB
constexpr int BWD_BARS = BWD_STAGE + 32768;
constexpr int BWD_SMEM = BWD_BARS + 128 + 1024;
constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DQ = 256, TM_DK = 384, TM_DV = 448;

// Descriptor = constant fields | (shared address >> 4): the issuing thread adds small address deltas instead of
// rebuilding the 64-bit word for every MMA (the single issuing thread is on the critical path of every tile).
__device__ __forceinline__ uint64_t desc_at(uint64_t fields, uint32_t addr) {
  return fields | static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
}

// geometry of (key block, query block) tile t as seen by one warp
struct BwdTile {
  int kb, qb, nkc, u_begin, u_end, q_warp0;
  bool warp_live;
  const float* bias_blk;  // first bias element of the warp's 32 rows at the key block's first column, or nullptr
};
__device__ __forceinline__ BwdTile bwd_tile(const AttnParams& p, int t, int nqb, int quad, int half, int h) {
  BwdTile g;
  g.kb = nqb == 2 ? (t >> 1) : t;
  g.qb = nqb == 2 ? (t & 1) : 0;
  g.nkc = (min(p.Lk - g.kb * 128, 128) + 15) >> 4;  // valid keys of the block in 16-column chunks
  const int nunit = (g.nkc + 1) >> 1;               // 32-column units, split between the two halves
  g.u_begin = half == 0 ? 0 : (nunit + 1) >> 1;
  g.u_end = half == 0 ? (nunit + 1) >> 1 : nunit;
  g.q_warp0 = g.qb * 128 + quad * 32;
  // a warp whose rows all lie beyond Lq has nothing to produce (its P/dS rows only feed dQ rows never stored)
  g.warp_live = g.q_warp0 < p.Lq;
  g.bias_blk = (p.bias && g.warp_live)
                   ? p.bias + h * p.bias_h_stride + static_cast<int64_t>(g.q_warp0) * p.bias_q_stride + g.kb * 128
                   : nullptr;
  return g;
}

// TMEM row (this thread's lane) -> 32 fp32 columns -> scaled bf16 -> the thread's row of a 128B-swizzled [128 x 64]
// staging tile (columns col0 .. col0+31), from where one TMA store writes the tile (rows past the sequence are clipped).
__device__ __forceinline__ void stage_row32(uint32_t tile_addr, int row, int col0, uint32_t tcol_addr, float mul) {
  uint32_t a[16], c[16];
  tmem_ld_32x16(tcol_addr, a);
  tmem_ld_32x16(tcol_addr + 16, c);
  tmem_wait_ld();
  st_shared_v4(tile_addr + swz_off(row, col0), pack_bf16x2(__uint_as_float(a[0]) * mul, __uint_as_float(a[1]) * mul),
               pack_bf16x2(__uint_as_float(a[2]) * mul, __uint_as_float(a[3]) * mul),
               pack_bf16x2(__uint_as_float(a[4]) * mul, __uint_as_float(a[5]) * mul),
               pack_bf16x2(__uint_as_float(a[6]) * mul, __uint_as_float(a[7]) * mul));
  st_shared_v4(tile_addr + swz_off(row, col0 + 8), pack_bf16x2(__uint_as_float(a[8]) * mul, __uint_as_float(a[9]) * mul),
               pack_bf16x2(__uint_as_float(a[10]) * mul, __uint_as_float(a[11]) * mul),
               pack_bf16x2(__uint_as_float(a[12]) * mul, __uint_as_float(a[13]) * mul),
               pack_bf16x2(__uint_as_float(a[14]) * mul, __uint_as_float(a[15]) * mul));
  st_shared_v4(tile_addr + swz_off(row, col0 + 16), pack_bf16x2(__uint_as_float(c[0]) * mul, __uint_as_float(c[1]) * mul),
               pack_bf16x2(__uint_as_float(c[2]) * mul, __uint_as_float(c[3]) * mul),
               pack_bf16x2(__uint_as_float(c[4]) * mul, __uint_as_float(c[5]) * mul),
               pack_bf16x2(__uint_as_float(c[6]) * mul, __uint_as_float(c[7]) * mul));
  st_shared_v4(tile_addr + swz_off(row, col0 + 24), pack_bf16x2(__uint_as_float(c[8]) * mul, __uint_as_float(c[9]) * mul),
               pack_bf16x2(__uint_as_float(c[10]) * mul, __uint_as_float(c[11]) * mul),
               pack_bf16x2(__uint_as_float(c[12]) * mul, __uint_as_float(c[13]) * mul),
               pack_bf16x2(__uint_as_float(c[14]) * mul, __uint_as_float(c[15]) * mul));
}

// delta = rowsum(dO ∘ O) of one query row and head, when no pre-kernel supplied it
__device__ __forceinline__ float row_delta(const AttnParams& p, int b, int q, int h) {
  const uint4* po = reinterpret_cast<const uint4*>(p.o + (static_cast<int64_t>(b) * p.Lq + q) * p.ld_o + h * 64);
  const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + (static_cast<int64_t>(b) * p.Lq + q) * p.ld_do + h * 64);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 a = __ldg(po + i), d = __ldg(pd + i);
    acc += bf16_lo(a.x) * bf16_lo(d.x) + bf16_hi(a.x) * bf16_hi(d.x) + bf16_lo(a.y) * bf16_lo(d.y) +
           bf16_hi(a.y) * bf16_hi(d.y) + bf16_lo(a.z) * bf16_lo(d.z) + bf16_hi(a.z) * bf16_hi(d.z) +
           bf16_lo(a.w) * bf16_lo(d.w) + bf16_hi(a.w) * bf16_hi(d.w);
  }
  return acc;
}

// Schedule of one CTA = one (b, h), up to 2 x 2 tiles of 128 queries x 128 keys, key block outer.  Three warpgroups:
// warps 0-7 compute (thread == query row, half == column range), warp 8 issues every MMA, warp 9 every bulk store
// (setmaxnreg moves the registers of warpgroup 2 to the compute warps).
//   * every operand tile (Q, dO of both query blocks, K, V of both key blocks) is TMA-loaded once, up front;
//   * tile t, compute warps: wait S/dP(t) (bar_s) -> P / dS INTO REGISTERS -> arrive "S/dP consumed" -> wait for the
//     previous tile's dQ / dK / dV chains (bar_pd: they read the P / dS tiles in shared memory) -> drain a finished key
//     block -> store the registers -> arrive "P/dS stored".  They never stop at a CTA-wide barrier;
//   * MMA warp: "consumed" -> S = Q·Kᵀ, dP = dO·Vᵀ of tile t+1;  "stored" -> dQ += dS·K, dK += dSᵀ·Q, dV += Pᵀ·dO of
//     tile t.  The CUDA-core pass of a tile overlaps the gradient chains of its predecessor although P / dS, S / dP
//     and the accumulators are all single-buffered (shared memory and TMEM are full: 224 KB, 512 columns);
//   * store warp: dS leaves by TMA store straight from the UMMA operand tile (rel-pos bias gradient); dK / dV of a
//     finished key block are staged as bf16 in that block's (now dead) K / V tiles, dQ at the end in the Q tiles, and
//     written by TMA stores whose 3-D maps clip the rows past the sequence.  Its elected lane joins bar_pd once the
//     dS store has read the tile, so nobody overwrites it early.
constexpr int BWD_THREADS = 384;
// named barriers (0 = __syncthreads): every one has a single arriving site (compute warps, bar.arrive) and a single waiting
// site (one issue warp, bar.sync), 256 + 32 threads — the MMA warp and the store warp wait on separate barriers
constexpr int NB_CONSUMED = 1, NB_STORED_MMA = 2, NB_FINAL_MMA = 3, NB_STORED_ST = 4, NB_FINAL_ST = 5;
constexpr int NB_COUNT = 256 + 32;

__device__ __forceinline__ void named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dk,
                const __grid_constant__ CUtensorMap tmap_dv, const __grid_constant__ CUtensorMap tmap_ds, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AT_TRACE(0);
  uint64_t* bar_ld = reinterpret_cast<uint64_t*>(smem + BWD_BARS);  // [0] Q0 dO0 K0 V0, [1] Q1 dO1, [2] K1 V1
  uint64_t* bar_s = bar_ld + 3;                                     // S / dP of a tile are in TMEM
  uint64_t* bar_pd = bar_ld + 4;  // dQ / dK / dV chains of a tile are complete (and its dS store has read the tile)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_ld + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Issue code runs warp-converged (warp index taken through a shuffle so that the compiler can prove it uniform) with
  // only the instruction itself under elect.sync: descriptors then live in uniform registers (tools/mma_bench.cu:
  // 58 / 67 cycles per N = 64 / 128 MMA instead of 91 from a divergent `threadIdx.x == 0` branch).
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int h = blockIdx.x, b = blockIdx.y;
  const int nqb = (p.Lq + 127) >> 7, nkb = (p.Lk + 127) >> 7;
  const int ntile = nqb * nkb;
  const uint32_t sbase = smem_u32(smem);

  if (warp_u == 8) {
    if (elect_one()) {
      const int kvb = p.kv_index ? p.kv_index[b] : b;
#pragma unroll
      for (int i = 0; i < 4; ++i) mbar_init(bar_ld + i, 1);
      mbar_init(bar_pd, p.ds_out ? 2 : 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(bar_ld, 4 * 16384);
      tma_load_2d(smem + BWD_SQ, &tmap_q, bar_ld, h * 64, b * p.Lq);
      tma_load_2d(smem + BWD_SK, &tmap_k, bar_ld, h * 64, kvb * p.Lk);
      tma_load_2d(smem + BWD_SDO, &tmap_do, bar_ld, h * 64, b * p.Lq);
      tma_load_2d(smem + BWD_SV, &tmap_v, bar_ld, h * 64, kvb * p.Lk);
      if (nqb > 1) {
        mbar_arrive_expect_tx(bar_ld + 1, 2 * 16384);
        tma_load_2d(smem + BWD_SQ + 16384, &tmap_q, bar_ld + 1, h * 64, b * p.Lq + 128);
        tma_load_2d(smem + BWD_SDO + 16384, &tmap_do, bar_ld + 1, h * 64, b * p.Lq + 128);
      }
      if (nkb > 1) {
        mbar_arrive_expect_tx(bar_ld + 2, 2 * 16384);
        tma_load_2d(smem + BWD_SK + 16384, &tmap_k, bar_ld + 2, h * 64, kvb * p.Lk + 128);
        tma_load_2d(smem + BWD_SV + 16384, &tmap_v, bar_ld + 2, h * 64, kvb * p.Lk + 128);
      }
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  AT_TRACE(1);

  if (warp_u >= 8) {
    // =========================== warpgroup 2: MMA issue (warp 8), bulk stores (warp 9) ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp_u == 8) {
      const uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t idesc_dkv = make_idesc_bf16(128, 64, 1, 1);
      const uint64_t f_k = make_smem_desc(0, 16, 1024);       // K-major operand
      const uint64_t f_mn = make_smem_desc(0, 8192, 1024);    // MN-major [k rows][64] operand (Q, dO, K as B)
      const uint64_t f_mnp = make_smem_desc(0, 16384, 1024);  // MN-major P / dS as A: 64-key chunks are 16 KB apart
      // S = Q[qb]·K[kb]ᵀ and dP = dO[qb]·V[kb]ᵀ, N = the block's valid keys rounded up to 16
      auto issue_scores = [&](int kb, int qb) {
        const int nkc = (min(p.Lk - kb * 128, 128) + 15) >> 4;
        const uint32_t idesc = make_idesc_bf16(128, nkc * 16, 0, 0);
        const uint64_t dq_ = desc_at(f_k, sbase + BWD_SQ + qb * 16384), dk_ = desc_at(f_k, sbase + BWD_SK + kb * 16384);
        const uint64_t ddo = desc_at(f_k, sbase + BWD_SDO + qb * 16384), dv_ = desc_at(f_k, sbase + BWD_SV + kb * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (elect_one()) umma_bf16(tmem + TM_S, dq_ + 2 * k, dk_ + 2 * k, idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (elect_one()) umma_bf16(tmem + TM_DP, ddo + 2 * k, dv_ + 2 * k, idesc, k != 0);
        if (elect_one()) umma_commit(bar_s);
        __syncwarp();  // the named barriers below are warp-aligned instructions: reconverge after the elected lane's work
      };
      mbar_wait(bar_ld, 0);
      tc_fence_after();
      issue_scores(0, 0);
      for (int t = 0; t < ntile; ++t) {
        const int kb = nqb == 2 ? (t >> 1) : t, qb = nqb == 2 ? (t & 1) : 0;
        const int nkc = (min(p.Lk - kb * 128, 128) + 15) >> 4;
        named_sync(NB_CONSUMED, NB_COUNT);  // every compute thread has read S / dP of tile t out of TMEM
        tc_fence_after();
        if (t + 1 < ntile) {
          const int kbn = nqb == 2 ? ((t + 1) >> 1) : t + 1, qbn = nqb == 2 ? ((t + 1) & 1) : 0;
          if (qbn == 1 && kbn == 0) mbar_wait(bar_ld + 1, 0);
          if (kbn == 1 && qbn == 0) mbar_wait(bar_ld + 2, 0);
          issue_scores(kbn, qbn);
        }
        AT_TRACE(46 + (t & 1) * 4);
        named_sync(NB_STORED_MMA, NB_COUNT);  // P / dS of tile t are in shared memory (fenced for the async proxy)
        tc_fence_after();
        AT_TRACE(47 + (t & 1) * 4);
        const uint32_t ads = sbase + BWD_SDS, ap = sbase + BWD_SP;
        const int nqc = (min(p.Lq - qb * 128, 128) + 15) >> 4;  // valid query rows in 16-row groups
        // dQ[qb] += dS · K          (A: dS K-major over keys; B: K tile MN-major, N = 64 dims)
        {
          const uint64_t db0 = desc_at(f_mn, sbase + BWD_SK + kb * 16384);
          for (int ks = 0; ks < nkc; ++ks) {
            const uint64_t da = desc_at(f_k, ads + (ks >> 2) * 16384 + (ks & 3) * 32);
            if (elect_one()) umma_bf16(tmem + TM_DQ + qb * 64, da, db0 + ks * 128, idesc_dq, (kb | ks) != 0);
          }
        }
        // dK += dSᵀ · Q[qb], dV += Pᵀ · dO[qb]   (A: dS / P MN-major over keys, K = query rows; B: Q / dO MN-major)
        {
          const uint64_t da0 = desc_at(f_mnp, ads), db0 = desc_at(f_mn, sbase + BWD_SQ + qb * 16384);
          for (int ks = 0; ks < nqc; ++ks)
            if (elect_one()) umma_bf16(tmem + TM_DK, da0 + ks * 128, db0 + ks * 128, idesc_dkv, (qb | ks) != 0);
        }
        {
          const uint64_t da0 = desc_at(f_mnp, ap), db0 = desc_at(f_mn, sbase + BWD_SDO + qb * 16384);
          for (int ks = 0; ks < nqc; ++ks)
            if (elect_one()) umma_bf16(tmem + TM_DV, da0 + ks * 128, db0 + ks * 128, idesc_dkv, (qb | ks) != 0);
        }
        if (elect_one()) umma_commit(bar_pd);
        __syncwarp();
        AT_TRACE(48 + (t & 1) * 4);
      }
      named_sync(NB_FINAL_MMA, NB_COUNT);  // every accumulator has been read out of TMEM
      tc_fence_after();
      tmem_dealloc(tmem, 512);
    } else if (warp_u == 9) {
      const bool lead = elect_one();
      int drain_kb = -1;
      for (int t = 0; t < ntile; ++t) {
        const int kb = nqb == 2 ? (t >> 1) : t, qb = nqb == 2 ? (t & 1) : 0;
        const int nkc = (min(p.Lk - kb * 128, 128) + 15) >> 4;
        named_sync(NB_STORED_ST, NB_COUNT);  // dS of tile t (and the key block staged during its pass) are in shared memory
        if (lead) {
          if (drain_kb >= 0) {
            tma_store_3d(&tmap_dk, sbase + BWD_SK + drain_kb * 16384, h * 64, drain_kb * 128, b);
            tma_store_3d(&tmap_dv, sbase + BWD_SV + drain_kb * 16384, h * 64, drain_kb * 128, b);
          }
          if (p.ds_out) {  // dS tile -> ds_out[b, h, q block, key block] (columns past pad16(Lk), rows past Lq clipped)
            for (int blk = 0; blk * 4 < nkc; ++blk)
              tma_store_4d(&tmap_ds, sbase + BWD_SDS + blk * 16384, kb * 128 + blk * 64, qb * 128, h, b);
            tma_store_commit();
            tma_store_wait_read();
            mbar_arrive(bar_pd);  // the dS tile may be overwritten (once the MMAs that read it have completed too)
          } else {
            tma_store_commit();
          }
        }
        __syncwarp();
        drain_kb = (qb == nqb - 1) ? kb : -1;
      }
      named_sync(NB_FINAL_ST, NB_COUNT);  // last key block's dK / dV and dQ are staged in the K / V / Q tiles
      if (lead) {
        tma_store_3d(&tmap_dk, sbase + BWD_SK + drain_kb * 16384, h * 64, drain_kb * 128, b);
        tma_store_3d(&tmap_dv, sbase + BWD_SV + drain_kb * 16384, h * 64, drain_kb * 128, b);
        for (int i = 0; i < nqb; ++i) tma_store_3d(&tmap_dq, sbase + BWD_SQ + i * 16384, h * 64, i * 128, b);
        tma_store_commit();
        tma_store_wait_read();  // shared memory must stay intact until the bulk stores have read it
      }
      __syncwarp();
    }
    return;
  }

  // ======================================= warpgroups 0, 1: compute ============================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const uint32_t stage_addr = sbase + BWD_STAGE + warp * 4096;

  // per-row statistics of both query blocks: lse (log2 domain) and delta = rowsum(dO ∘ O)
  const int64_t stat0 = (static_cast<int64_t>(b) * p.H + h) * p.Lq;
  float lse_0 = 0.f, lse_1 = 0.f, dl_0 = 0.f, dl_1 = 0.f;
  if (row < p.Lq) {
    lse_0 = p.lse[stat0 + row];
    dl_0 = p.delta ? __ldg(p.delta + stat0 + row) : row_delta(p, b, row, h);
  }
  if (128 + row < p.Lq) {
    lse_1 = p.lse[stat0 + 128 + row];
    dl_1 = p.delta ? __ldg(p.delta + stat0 + 128 + row) : row_delta(p, b, 128 + row, h);
  }
  const DropCfg dc = make_drop(p.dropout_p);
  const uint64_t doff = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
  uint32_t s_phase = 0, pd_phase = 0;

  BwdTile g = bwd_tile(p, 0, nqb, quad, half, h);
  float4 bv[8];
  if (g.bias_blk && g.u_begin < g.u_end) bias_issue(g.bias_blk + g.u_begin * 32, p.bias_q_stride, p.Lq - g.q_warp0, lane, bv);
  AT_TRACE(2);

  int drain_kb = -1;  // key block whose dK / dV are final with the pending bar_pd phase and not yet staged
  for (int t = 0; t < ntile; ++t) {
    const int kb = g.kb, qb = g.qb, nkc = g.nkc;
    mbar_wait_spin_warp(bar_s, s_phase);
    s_phase ^= 1;
    tc_fence_after();
    AT_TRACE(4 + t * 6);

    // ---- P and dS of this (q block, key block) tile into registers ----
    const int q = qb * 128 + row;
    const bool qvalid = q < p.Lq;
    const float* mask_row = (p.mask && qvalid) ? p.mask + b * p.mask_b_stride + q * p.mask_q_stride : nullptr;
    const uint64_t drop_base = (static_cast<uint64_t>(b * p.H + h) * p.Lq + q) * p.Lk_pad;
    const float my_lse = qb ? lse_1 : lse_0, my_delta = qb ? dl_1 : dl_0;
    uint32_t oP[4][8], oD[4][8];
#pragma unroll
    for (int ui = 0; ui < 2; ++ui) {
      const int u = g.u_begin + ui;
      if (g.warp_live && u < g.u_end) {  // warp-uniform
        if (g.bias_blk) {
          bias_stage(bv, stage_addr, lane);
          if (u + 1 < g.u_end) bias_issue(g.bias_blk + (u + 1) * 32, p.bias_q_stride, p.Lq - g.q_warp0, lane, bv);
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int c = 2 * u + v;
          if (c < nkc) {  // warp-uniform
            const int k0 = kb * 128 + c * 16;
            uint32_t s[16], dp[16];
            tmem_ld_32x16(trow + TM_S + c * 16, s);
            tmem_ld_32x16(trow + TM_DP + c * 16, dp);
            float add[16];
            if (g.bias_blk) {
              bias_read16(stage_addr, lane, v, add);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) add[j] = 0.f;
            }
            if (mask_row) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(mask_row + k0 + j));
                add[j] += m.x; add[j + 1] += m.y; add[j + 2] += m.z; add[j + 3] += m.w;
              }
            }
            tmem_wait_ld();  // .sync.aligned: reached by the whole warp, never inside a divergent branch
            float pr[16], ds[16];
            if (qvalid) {
              // P = 2^(s·scale·log2e + (bias + mask)·log2e − lse): two FMAs and one MUFU per element
              if (k0 + 16 <= p.Lk) {  // warp-uniform: only the sequence's last chunk needs the key bound
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  pr[j] = fast_exp2(fmaf(__uint_as_float(s[j]), p.scale_log2, fmaf(add[j], kLog2e, -my_lse)));
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  pr[j] = (k0 + j < p.Lk)
                              ? fast_exp2(fmaf(__uint_as_float(s[j]), p.scale_log2, fmaf(add[j], kLog2e, -my_lse)))
                              : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) ds[j] = __uint_as_float(dp[j]);
              if (p.dropout_p > 0.f) {
#pragma unroll
                for (int j = 0; j < 16; j += 8) {
                  float k[8];
                  drop8(p.seed, doff, (drop_base + k0 + j) >> 3, dc, k);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    // dS uses the un-dropped P; the P that feeds dV is the dropped one
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
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              oP[ui * 2 + v][j] = pack_bf16x2(pr[2 * j], pr[2 * j + 1]);
              oD[ui * 2 + v][j] = pack_bf16x2(ds[2 * j], ds[2 * j + 1]);
            }
          }
        }
        if (g.bias_blk) __syncwarp();  // the staging tile is rewritten by the next unit
      }
    }
    tc_fence_before();
    named_arrive(NB_CONSUMED, NB_COUNT);  // S / dP may be overwritten by the next tile's
    AT_TRACE(5 + t * 6);

    // the next tile's first bias block: requested now, consumed after the wait for its S / dP
    const bool last = t == ntile - 1;
    BwdTile gn = g;
    if (!last) {
      gn = bwd_tile(p, t + 1, nqb, quad, half, h);
      if (gn.bias_blk && gn.u_begin < gn.u_end)
        bias_issue(gn.bias_blk + gn.u_begin * 32, p.bias_q_stride, p.Lq - gn.q_warp0, lane, bv);
    }

    // ---- the previous tile's MMAs have read P / dS (and its dS store too): drain a finished key block, store ----
    if (t > 0) {
      mbar_wait_spin_warp(bar_pd, pd_phase);
      pd_phase ^= 1;
      tc_fence_after();
      if (drain_kb >= 0) {  // dK / dV of the finished key block -> bf16 staging in its K / V tiles (dead from here on)
        stage_row32(sbase + BWD_SK + drain_kb * 16384, row, half * 32, trow + TM_DK + half * 32, p.scale);
        stage_row32(sbase + BWD_SV + drain_kb * 16384, row, half * 32, trow + TM_DV + half * 32, 1.0f);
      }
    }
#pragma unroll
    for (int ui = 0; ui < 2; ++ui) {
      const int u = g.u_begin + ui;
      if (g.warp_live && u < g.u_end) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int c = 2 * u + v;
          if (c < nkc) {
            const uint32_t o0 = swz_off(row, c * 16), o1 = swz_off(row, c * 16 + 8);
            const uint32_t(&a)[8] = oP[ui * 2 + v];
            const uint32_t(&d)[8] = oD[ui * 2 + v];
            st_shared_v4(sbase + BWD_SP + o0, a[0], a[1], a[2], a[3]);
            st_shared_v4(sbase + BWD_SP + o1, a[4], a[5], a[6], a[7]);
            st_shared_v4(sbase + BWD_SDS + o0, d[0], d[1], d[2], d[3]);
            st_shared_v4(sbase + BWD_SDS + o1, d[4], d[5], d[6], d[7]);
          }
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_arrive(NB_STORED_MMA, NB_COUNT);
    named_arrive(NB_STORED_ST, NB_COUNT);
    AT_TRACE(6 + t * 6);
    drain_kb = (qb == nqb - 1) ? kb : -1;
    g = gn;
  }

  // ---- last key block's dK / dV and dQ: staged in the K / V / Q tiles (every MMA has completed), TMA-stored ----
  mbar_wait_spin_warp(bar_pd, pd_phase);
  tc_fence_after();
  AT_TRACE(40);
  stage_row32(sbase + BWD_SK + drain_kb * 16384, row, half * 32, trow + TM_DK + half * 32, p.scale);
  stage_row32(sbase + BWD_SV + drain_kb * 16384, row, half * 32, trow + TM_DV + half * 32, 1.0f);
  for (int i = 0; i < nqb; ++i)
    stage_row32(sbase + BWD_SQ + i * 16384, row, half * 32, trow + TM_DQ + i * 64 + half * 32, p.scale);
  AT_TRACE(41);
  fence_proxy_async_smem();
  tc_fence_before();
  named_arrive(NB_FINAL_MMA, NB_COUNT);
  named_arrive(NB_FINAL_ST, NB_COUNT);
  AT_TRACE(60);
}

// ---------------------------------------------------------------------------------------------
// relative-position bias gather / scatter (HBM-bound, tiny)
// ---------------------------------------------------------------------------------------------
__global__ void relpos_gather_kernel(const float* __restrict__ table, const int64_t* __restrict__ index, int N, int H,
                                     float* __restrict__ out, int64_t ld_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // key (column, padded)
  const int i = blockIdx.y, h = blockIdx.z;
  if (j >= ld_out) return;
  float v = 0.f;
  if (j < N) v = __ldg(table + __ldg(index + static_cast<int64_t>(i) * N + j) * H + h);
  out[(static_cast<int64_t>(h) * N + i) * ld_out + j] = v;
}

__global__ void relpos_scatter_kernel(const __nv_bfloat16* __restrict__ ds, int B, int H, int N, int64_t sb, int64_t sh,
                                      int64_t sq, const int64_t* __restrict__ index, float* __restrict__ dtable) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y, h = blockIdx.z;
  if (j >= N) return;
  const __nv_bfloat16* src = ds + h * sh + i * sq + j;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += __bfloat162float(src[b * sb]);
  atomicAdd(dtable + __ldg(index + static_cast<int64_t>(i) * N + j) * H + h, s);
}


// ---------------------------------------------------------------------------------------------
// Attention maps for the callers that ask for them (output_attentions: knowledge distillation; save_attention: Grad-CAM).
// The fused kernels never materialise [B,H,Lq,Lk]; this streaming kernel rebuilds it from Q, K and the forward's
// log-sum-exp:  mode 0: P = softmax(scale·QKᵀ + bias + mask) (pre-dropout, what the reference returns, models/beit2.py:152,
// models/xbert.py:392-410);  mode 1: dL/dP = (dO·Vᵀ) ∘ dropout keep-scale (the gradient models/xbert.py:394-396 hooks).
// Block = 32 query rows x 64 keys of one (b, h); thread = one row x 8 keys; HBM-bound on the fp32 output.
// ---------------------------------------------------------------------------------------------
constexpr int PB_ROWS = 32, PB_KEYS = 64, PB_LD = 66;  // smem rows padded to 33 words: 2-way conflicts at worst

__global__ void __launch_bounds__(256)
attn_probs_kernel(const __nv_bfloat16* __restrict__ a, int64_t ld_a, const __nv_bfloat16* __restrict__ bm, int64_t ld_b,
                  const AttnParams p, int mode, float* __restrict__ out) {
  __shared__ __nv_bfloat16 sA[PB_ROWS * PB_LD];
  __shared__ __nv_bfloat16 sB[PB_KEYS * PB_LD];
  const int bh = blockIdx.z, b = bh / p.H, h = bh - b * p.H;
  const int q0 = blockIdx.y * PB_ROWS, k0 = blockIdx.x * PB_KEYS;
  const int kvb = p.kv_index ? p.kv_index[b] : b;
  for (int i = threadIdx.x; i < PB_ROWS * 32; i += 256) {  // 32 rows x 32 bf16-pairs
    const int r = i >> 5, c = (i & 31) * 2;
    uint32_t v = 0u;
    if (q0 + r < p.Lq) v = *reinterpret_cast<const uint32_t*>(a + (static_cast<int64_t>(b) * p.Lq + q0 + r) * ld_a + h * 64 + c);
    *reinterpret_cast<uint32_t*>(sA + r * PB_LD + c) = v;
  }
  for (int i = threadIdx.x; i < PB_KEYS * 32; i += 256) {
    const int r = i >> 5, c = (i & 31) * 2;
    uint32_t v = 0u;
    if (k0 + r < p.Lk) v = *reinterpret_cast<const uint32_t*>(bm + (static_cast<int64_t>(kvb) * p.Lk + k0 + r) * ld_b + h * 64 + c);
    *reinterpret_cast<uint32_t*>(sB + r * PB_LD + c) = v;
  }
  __syncthreads();
  const int r = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 8;
  const int q = q0 + r;
  if (q >= p.Lq) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int d = 0; d < 64; d += 2) {
    const uint32_t av = *reinterpret_cast<const uint32_t*>(sA + r * PB_LD + d);
    const float a0 = bf16_lo(av), a1 = bf16_hi(av);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t bv = *reinterpret_cast<const uint32_t*>(sB + (c0 + j) * PB_LD + d);
      acc[j] = fmaf(a0, bf16_lo(bv), fmaf(a1, bf16_hi(bv), acc[j]));
    }
  }
  float* dst = out + ((static_cast<int64_t>(bh) * p.Lq + q) * p.Lk);
  if (mode == 0) {
    const float lse = p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Lq + q];
    const float* bias_row = p.bias ? p.bias + h * p.bias_h_stride + static_cast<int64_t>(q) * p.bias_q_stride : nullptr;
    const float* mask_row = p.mask ? p.mask + b * p.mask_b_stride + q * p.mask_q_stride : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + c0 + j;
      if (k < p.Lk) {
        float add = 0.f;
        if (bias_row) add += __ldg(bias_row + k);
        if (mask_row) add += __ldg(mask_row + k);
        dst[k] = fast_exp2(fmaf(acc[j], p.scale_log2, add * kLog2e) - lse);
      }
    }
  } else {
    float keep[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (p.dropout_p > 0.f) {
      const DropCfg dc = make_drop(p.dropout_p);
      const uint64_t doff = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
      const uint64_t base = (static_cast<uint64_t>(b * p.H + h) * p.Lq + q) * p.Lk_pad + k0 + c0;  // multiple of 8
      drop8(p.seed, doff, base >> 3, dc, keep);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + c0 + j;
      if (k < p.Lk) dst[k] = acc[j] * keep[j];
    }
  }
}

}  // namespace
}  // namespace x2k

using namespace x2k;

extern "C" int x2k_attn_fwd(const X2kAttnArgs* args, void* stream_) {
  X2K_REQUIRE(args != nullptr, "x2k_attn_fwd: args is NULL");
  const X2kAttnArgs& a = *args;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_common(a, "x2k_attn_fwd")) return rc;
  {
    const int rc = attn_pack_fwd(a, stream);  // short query sequences: several sequences per MMA tile
    if (rc <= 0) return rc;
  }
  X2K_REQUIRE(a.kv_groups == nullptr, "x2k_attn_fwd: kv_groups given but the shape is not eligible for the grouped kernel");
  static const bool force_long = getenv("X2K_ATTN_LONG") != nullptr;  // developer switch: time the key-blocked kernels on short shapes
  if (a.Lk > 256 || force_long) return attn_long_fwd(a, stream);  // key-blocked online-softmax kernel (attn_long.cu)
  AttnParams p;
  fill_params(a, p);
  const int n_kv = a.n_kv > 0 ? a.n_kv : a.B;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_q, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, a.k, static_cast<uint64_t>(n_kv) * a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_k, p.Lk_pad, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, a.v, static_cast<uint64_t>(n_kv) * a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_v, p.Lk_pad, 64))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr_set = true;
  }
  dim3 grid((a.Lq + 127) / 128, a.H, a.B);
  attn_fwd_kernel<<<grid, ATT_THREADS, FWD_SMEM, stream>>>(tq, tk, tv, p);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

namespace x2k {
// delta[b,h,i] = sum_d O[b,i,h,d] * dO[b,i,h,d]: one warp per token row, lane -> 16-byte chunk (8 bf16) of the row,
// 8 consecutive lanes = one head, reduced with three xor-shuffles.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t ld_o, const __nv_bfloat16* __restrict__ d_o, int64_t ld_do,
                  int rows, int H, int Lq, float* __restrict__ delta) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / Lq, i = row - b * Lq;
  const uint4* po = reinterpret_cast<const uint4*>(o + static_cast<int64_t>(row) * ld_o);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + static_cast<int64_t>(row) * ld_do);
  for (int c0 = 0; c0 < H * 8; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    if (c < H * 8) {
      const uint4 a = __ldg(po + c), d = __ldg(pd + c);
      acc = bf16_lo(a.x) * bf16_lo(d.x) + bf16_hi(a.x) * bf16_hi(d.x) + bf16_lo(a.y) * bf16_lo(d.y) + bf16_hi(a.y) * bf16_hi(d.y) +
            bf16_lo(a.z) * bf16_lo(d.z) + bf16_hi(a.z) * bf16_hi(d.z) + bf16_lo(a.w) * bf16_lo(d.w) + bf16_hi(a.w) * bf16_hi(d.w);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if ((lane & 7) == 0 && c < H * 8) delta[(static_cast<int64_t>(b) * H + (c >> 3)) * Lq + i] = acc;
  }
}

int attn_delta_launch(const X2kAttnArgs& a, cudaStream_t stream) {
  const int rows = a.B * a.Lq;
  attn_delta_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(a.o), a.ld_o,
                                                        static_cast<const __nv_bfloat16*>(a.d_o), a.ld_do, rows, a.H, a.Lq,
                                                        a.delta_ws);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}
}  // namespace x2k

extern "C" int x2k_attn_bwd(const X2kAttnArgs* args, void* stream_) {
  X2K_REQUIRE(args != nullptr, "x2k_attn_bwd: args is NULL");
  const X2kAttnArgs& a = *args;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = check_common(a, "x2k_attn_bwd")) return rc;
  X2K_REQUIRE(a.d_o && a.dq && a.dk && a.dv, "x2k_attn_bwd: NULL d_o/dq/dk/dv");
  X2K_REQUIRE(a.ld_do % 8 == 0 && a.ld_dq % 8 == 0 && a.ld_dk % 8 == 0 && a.ld_dv % 8 == 0, "x2k_attn_bwd: ld alignment");
  if (a.delta_ws) {
    X2K_REQUIRE(a.ld_o % 8 == 0 && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.d_o) & 15) == 0,
                "x2k_attn_bwd: o / d_o must be 16-byte aligned rows for the delta pre-kernel");
    if (int rc = attn_delta_launch(a, stream)) return rc;
  }
  {
    const int rc = attn_pack_bwd(a, stream);
    if (rc <= 0) return rc;
  }
  X2K_REQUIRE(a.kv_groups == nullptr, "x2k_attn_bwd: kv_groups given but the shape is not eligible for the grouped kernel");
  static const bool force_long = getenv("X2K_ATTN_LONG") != nullptr;
  if (a.Lq > 256 || a.Lk > 256 || (force_long && a.dq_ws)) return attn_long_bwd(a, stream);  // key-blocked kernel, dQ through an fp32 workspace
  X2K_REQUIRE(!a.ds_out || (a.ds_q_stride % 8 == 0 && a.ds_h_stride % 8 == 0 && a.ds_b_stride % 8 == 0 &&
                            a.ds_q_stride >= ((a.Lk + 15) & ~15)),
              "x2k_attn_bwd: ds_out strides must be multiples of 8 and cover Lk_pad");
  AttnParams p;
  fill_params(a, p);
  const int n_kv = a.n_kv > 0 ? a.n_kv : a.B;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  X2K_REQUIRE(al16(a.d_o) && al16(a.dq) && al16(a.dk) && al16(a.dv) && al16(a.ds_out), "x2k_attn_bwd: 16-byte alignment");
  CUtensorMap tq, tk, tv, tdo, tdq, tdk, tdv, tds;
  int rc;
  const uint64_t cols = static_cast<uint64_t>(a.H) * 64;
  if ((rc = make_tmap_bf16_2d(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, cols, a.ld_q, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, a.k, static_cast<uint64_t>(n_kv) * a.Lk, cols, a.ld_k, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, a.v, static_cast<uint64_t>(n_kv) * a.Lk, cols, a.ld_v, 128, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, a.d_o, static_cast<uint64_t>(a.B) * a.Lq, cols, a.ld_do, 128, 64))) return rc;
  // gradients leave as TMA stores of [128 x 64] tiles; the per-sequence view clips the rows past Lq / Lk
  if ((rc = make_tmap_bf16_seq3d(&tdq, a.dq, a.B, a.Lq, cols, a.ld_dq, 128))) return rc;
  if ((rc = make_tmap_bf16_seq3d(&tdk, a.dk, a.B, a.Lk, cols, a.ld_dk, 128))) return rc;
  if ((rc = make_tmap_bf16_seq3d(&tdv, a.dv, a.B, a.Lk, cols, a.ld_dv, 128))) return rc;
  if (a.ds_out) {
    if ((rc = make_tmap_bf16_4d(&tds, a.ds_out, p.Lk_pad, a.Lq, a.H, a.B, a.ds_q_stride, a.ds_h_stride, a.ds_b_stride))) return rc;
  } else {
    tds = tdq;  // never used by the kernel
  }
  static bool attr_set = false;
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr_set = true;
  }
  dim3 grid(a.H, a.B);
  attn_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, tds, p);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int64_t x2k_attn_bwd_workspace_bytes(const X2kAttnArgs* args) {
  if (args == nullptr) return 0;
  const X2kAttnArgs& a = *args;
  if (a.B <= 0 || a.H <= 0 || a.Lq <= 0 || a.Lk <= 0) return 0;
  // the packed short-sequence kernels (Lq <= 64 and Lk <= 256) and the whole-range kernels (both <= 256) need none
  static const bool force_long = getenv("X2K_ATTN_LONG") != nullptr;  // developer switch (see x2k_attn_bwd)
  return (a.Lq > 256 || a.Lk > 256 || (force_long && a.Lq > 64)) ? attn_long_bwd_ws_bytes(a) : 0;
}

extern "C" int x2k_attn_probs(const X2kAttnArgs* args, int32_t mode, float* out, void* stream_) {
  X2K_REQUIRE(args != nullptr && out != nullptr, "x2k_attn_probs: NULL argument");
  const X2kAttnArgs& a = *args;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(mode == 0 || mode == 1, "x2k_attn_probs: mode must be 0 (probabilities) or 1 (their gradient)");
  X2K_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "x2k_attn_probs: bad shape");
  X2K_REQUIRE(static_cast<int64_t>(a.B) * a.H <= 65535, "x2k_attn_probs: B*H > 65535");
  AttnParams p;
  fill_params(a, p);
  const void *pa, *pb;
  int64_t lda, ldb;
  if (mode == 0) {
    X2K_REQUIRE(a.q && a.k && a.lse, "x2k_attn_probs: mode 0 needs q, k and the forward's lse");
    pa = a.q; lda = a.ld_q; pb = a.k; ldb = a.ld_k;
  } else {
    X2K_REQUIRE(a.d_o && a.v, "x2k_attn_probs: mode 1 needs d_o and v");
    pa = a.d_o; lda = a.ld_do; pb = a.v; ldb = a.ld_v;
  }
  X2K_REQUIRE(lda % 2 == 0 && ldb % 2 == 0, "x2k_attn_probs: leading dimensions must be even");
  dim3 grid((a.Lk + PB_KEYS - 1) / PB_KEYS, (a.Lq + PB_ROWS - 1) / PB_ROWS, a.B * a.H);
  attn_probs_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pa), lda, static_cast<const __nv_bfloat16*>(pb), ldb,
                                              p, mode, out);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int x2k_relpos_bias_gather(const float* table, const int64_t* index, int32_t N, int32_t H, float* out,
                                      int64_t ld_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(table && index && out && N > 0 && H > 0 && ld_out >= N, "x2k_relpos_bias_gather: bad arguments");
  dim3 grid((ld_out + 127) / 128, N, H);
  relpos_gather_kernel<<<grid, 128, 0, stream>>>(table, index, N, H, out, ld_out);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int x2k_relpos_bias_scatter(const void* ds_bf16, int32_t B, int32_t H, int32_t N, int64_t ds_b_stride,
                                       int64_t ds_h_stride, int64_t ds_q_stride, const int64_t* index, float* dtable,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(ds_bf16 && index && dtable && B > 0 && H > 0 && N > 0, "x2k_relpos_bias_scatter: bad arguments");
  dim3 grid((N + 127) / 128, N, H);
  relpos_scatter_kernel<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(ds_bf16), B, H, N, ds_b_stride,
                                                  ds_h_stride, ds_q_stride, index, dtable);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

#ifdef X2K_ATTN_TRACE
extern "C" int x2k_debug_attn_fwd_trace(long long* out128) {
  return cudaMemcpyFromSymbol(out128, g_attn_fwd_trace, sizeof(long long) * 128) == cudaSuccess ? 0 : -2;
}
extern "C" int x2k_debug_attn_trace(long long* out128) {
  return cudaMemcpyFromSymbol(out128, g_attn_trace, sizeof(long long) * 192) == cudaSuccess ? 0 : -2;
}
#endif
