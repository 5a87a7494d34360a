// attn_pack.cu — packed attention for SHORT query sequences (Lq <= 64, head_dim 64) on tcgen05, sm_100a.
//
// attn.cu gives one (sequence, head) a whole 128-row MMA tile; with the 40-token captions of X²-VLM two
// thirds of every tile are padding and each CTA is a chain of dependent latencies (TMA -> MMA -> softmax ->
// MMA -> store) with almost no work behind it.  Here several sequences share ONE tile:
//   * "self" mode  (BERT text self-attention, models/xbert.py:364-410): G = floor(128 / Lq8) consecutive
//     sequences are stacked as the rows of the Q tile and as the key columns of the K tile; row r only looks
//     at the key columns of its own sequence (a block-diagonal score matrix, the off-diagonal blocks are
//     written as P = 0), so S, P·V, dQ, dK, dV are each ONE MMA chain for G sequences;
//   * "cross" mode (fusion layers, text queries over image keys, models/xbert.py:343-349): the sequences
//     that look at the SAME image K/V (kv_index) are grouped by x2k_attn_group_build; a work item stacks up
//     to G of them over the one shared K/V tile.  The backward walks all items of a K/V source inside one
//     CTA and keeps dK/dV in TMEM across them, so dK/dV leave the SM once per source, already summed over
//     the query sequences (no per-sequence dK/dV round trip through HBM, no segment-sum pass).
// Slots are 8-row aligned (Lq8 = Lq rounded up to 8) so every slot is a whole number of 128B-swizzle atoms
// and is fetched by its own TMA box; unused slots are fetched from an out-of-bounds coordinate (zero fill).
// The dropout stream is indexed exactly like attn.cu's (element ((b*H+h)*Lq + i)*Lk_pad + j).
#include <cstdlib>

#include "common.cuh"

namespace x2k {
namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int PK_THREADS = 256;
constexpr int PK_MAX_G = 8;
constexpr int PK_ITEM_INTS = 12;  // {src, n, b[0..7], -, -}
constexpr int PK_HDR_INTS = 4;    // {n_items, G, n_kv, B}

// Optional phase trace (compile with -DX2K_PACK_TRACE): clock64 stamps of one CTA, read back by x2k_debug_pack_trace.
#ifdef X2K_PACK_TRACE
__device__ long long g_pack_trace[2][64];
#define PK_TRACE(i)                                                                           \
  do {                                                                                        \
    if (blockIdx.x == 40 && blockIdx.y == 3 && (threadIdx.x == 0 || threadIdx.x == 200))      \
      g_pack_trace[threadIdx.x ? 1 : 0][(i)] = clock64();                                     \
  } while (0)
#else
#define PK_TRACE(i) do {} while (0)
#endif

struct PackParams {
  const float* delta;  // [B,H,Lq] rowsum(dO ∘ O) from the pre-kernel, or nullptr
  int B, H, Lq, Lk, Lq8, Lk8, G;
  int N;        // key columns of the tile, multiple of 16 (self: pad16(G*Lk8), cross: pad16(Lk))
  int Lk_pad;   // pad16(Lk): row stride of the dropout element index (same stream as attn.cu)
  int cross, n_kv;
  const int32_t* table;
  int items_off;  // int offset of the item records inside `table`
  float scale_log2, scale;
  const float* mask;
  int64_t mask_b_stride, mask_q_stride;
  float dropout_p;
  uint64_t seed, offset;
  const uint64_t* offset_dev;
  __nv_bfloat16* o;
  int64_t ld_o;
  float* lse;
  const __nv_bfloat16* d_o;
  int64_t ld_do;
  __nv_bfloat16 *dq, *dk, *dv;
  int64_t ld_dq, ld_dk, ld_dv;
  // shared-memory byte offsets (from the 1024-aligned base)
  uint32_t off_do, off_k, off_v, off_p, off_ds, off_red, off_bar;
  uint32_t tmem_cols;
  uint32_t tm_dk, tm_dv;  // backward: first TMEM column of the dK / dV accumulators
};

// byte offset of the 16-byte chunk holding elements [k0, k0+8) of row `row` inside a K-major, 128B-swizzled tile
// made of [128 rows x 64 elements] blocks (block stride 16 KB)
__device__ __forceinline__ uint32_t swz_off(int row, int k0) {
  const int blk = k0 >> 6, c = (k0 & 63) >> 3;
  return blk * 16384 + row * 128 + ((c ^ (row & 7)) << 4);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void zero_smem(uint32_t addr, int bytes, int tid, int nthreads) {
  for (int i = tid * 16; i < bytes; i += nthreads * 16) st_shared_v4(addr + i, 0u, 0u, 0u, 0u);
}

// sequence id sitting in `slot` of work item `item`, or -1
__device__ __forceinline__ int item_seq(const PackParams& p, int item, int slot) {
  if (slot >= p.G) return -1;
  if (!p.cross) {
    const int b = item * p.G + slot;
    return b < p.B ? b : -1;
  }
  return __ldg(p.table + p.items_off + item * PK_ITEM_INTS + 2 + slot);
}

// 16 fp32 values -> scaled bf16 -> 32 bytes of global memory
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const uint32_t (&s)[16], float mul) {
  uint4 v;
  v.x = pack_bf16x2(__uint_as_float(s[0]) * mul, __uint_as_float(s[1]) * mul);
  v.y = pack_bf16x2(__uint_as_float(s[2]) * mul, __uint_as_float(s[3]) * mul);
  v.z = pack_bf16x2(__uint_as_float(s[4]) * mul, __uint_as_float(s[5]) * mul);
  v.w = pack_bf16x2(__uint_as_float(s[6]) * mul, __uint_as_float(s[7]) * mul);
  *reinterpret_cast<uint4*>(dst) = v;
  v.x = pack_bf16x2(__uint_as_float(s[8]) * mul, __uint_as_float(s[9]) * mul);
  v.y = pack_bf16x2(__uint_as_float(s[10]) * mul, __uint_as_float(s[11]) * mul);
  v.z = pack_bf16x2(__uint_as_float(s[12]) * mul, __uint_as_float(s[13]) * mul);
  v.w = pack_bf16x2(__uint_as_float(s[14]) * mul, __uint_as_float(s[15]) * mul);
  *reinterpret_cast<uint4*>(dst + 8) = v;
}
// TMEM lane (this thread's row) -> 64 / NPART fp32 columns starting at part * (64 / NPART) -> scaled bf16 -> global
template <int NPART>
__device__ __forceinline__ void store_row_part(__nv_bfloat16* dst_row, uint32_t tcol_addr, int part, float mul, bool valid) {
  constexpr int W = 64 / NPART;  // 32 or 16 columns
  uint32_t a[16], b[16];
  tmem_ld_32x16(tcol_addr + part * W, a);
  if (W == 32) tmem_ld_32x16(tcol_addr + part * W + 16, b);
  tmem_wait_ld();
  if (valid) {
    store16_bf16(dst_row + part * W, a, mul);
    if (W == 32) store16_bf16(dst_row + part * W + 16, b, mul);
  }
}

// TMEM lane (this thread's row) -> 64 / NPART fp32 columns starting at part * (64 / NPART) -> scaled bf16 -> the row's
// 16-byte chunks of a 128B-swizzled [rows x 64] staging tile (rows 128 B apart, tile base 1024-aligned): what a TMA
// store with SWIZZLE_128B expects.  The gradients leave the SM as whole tiles instead of 32 scattered 16-byte stores
// per warp instruction (the drains were a quarter of the backward CTA's life, tools/trace_attn.py).
template <int NPART>
__device__ __forceinline__ void stage_row_part(uint32_t tile_addr, int row, uint32_t tcol_addr, int part, float mul,
                                               bool valid) {
  constexpr int W = 64 / NPART;  // 32 or 16 columns
  uint32_t a[16], b[16];
  tmem_ld_32x16(tcol_addr + part * W, a);
  if (W == 32) tmem_ld_32x16(tcol_addr + part * W + 16, b);
  tmem_wait_ld();
  if (!valid) return;  // the row lies outside the staging tile (or must keep its zero padding)
  const uint32_t base = tile_addr + row * 128;
  const int c0 = (part * W) >> 3;  // first 16-byte chunk of this part
#pragma unroll
  for (int i = 0; i < 2; ++i)
    st_shared_v4(base + (((c0 + i) ^ (row & 7)) << 4), pack_bf16x2(__uint_as_float(a[8 * i]) * mul, __uint_as_float(a[8 * i + 1]) * mul),
                 pack_bf16x2(__uint_as_float(a[8 * i + 2]) * mul, __uint_as_float(a[8 * i + 3]) * mul),
                 pack_bf16x2(__uint_as_float(a[8 * i + 4]) * mul, __uint_as_float(a[8 * i + 5]) * mul),
                 pack_bf16x2(__uint_as_float(a[8 * i + 6]) * mul, __uint_as_float(a[8 * i + 7]) * mul));
  if (W == 32) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      st_shared_v4(base + (((c0 + 2 + i) ^ (row & 7)) << 4),
                   pack_bf16x2(__uint_as_float(b[8 * i]) * mul, __uint_as_float(b[8 * i + 1]) * mul),
                   pack_bf16x2(__uint_as_float(b[8 * i + 2]) * mul, __uint_as_float(b[8 * i + 3]) * mul),
                   pack_bf16x2(__uint_as_float(b[8 * i + 4]) * mul, __uint_as_float(b[8 * i + 5]) * mul),
                   pack_bf16x2(__uint_as_float(b[8 * i + 6]) * mul, __uint_as_float(b[8 * i + 7]) * mul));
  }
}

// additive mask (already in the log2 domain) of the 16 key columns [col0, col0+16) of this thread's row;
// kk0 = col0 - klo is the key index of the first column inside the row's own key segment (multiple of 4).
__device__ __forceinline__ void load_mask16(const float* mask_row, int kk0, int Lk, float (&add)[16]) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const int kk = kk0 + j;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mask_row != nullptr && kk >= 0 && kk < Lk) m = __ldg(reinterpret_cast<const float4*>(mask_row + kk));
    add[j] = m.x * kLog2e; add[j + 1] = m.y * kLog2e; add[j + 2] = m.z * kLog2e; add[j + 3] = m.w * kLog2e;
  }
}

// Everything a thread needs to know about the query row it owns (thread == TMEM lane == tile row).
struct RowInfo {
  int b, q;       // sequence, query index (valid only if ok)
  bool ok;        // the row holds a real query
  int klo;        // first key column of the row's key segment
  const float* mask_row;
  uint64_t drop_base;
};
__device__ __forceinline__ RowInfo row_info(const PackParams& p, int item, int row, int h) {
  RowInfo r;
  const int g = row / p.Lq8;
  r.q = row - g * p.Lq8;
  r.b = item_seq(p, item, g);
  r.ok = r.b >= 0 && r.q < p.Lq;
  r.klo = p.cross ? 0 : g * p.Lk8;
  r.mask_row = (p.mask != nullptr && r.ok) ? p.mask + r.b * p.mask_b_stride + r.q * p.mask_q_stride : nullptr;
  r.drop_base = r.ok ? (static_cast<uint64_t>(r.b * p.H + h) * p.Lq + r.q) * p.Lk_pad : 0;
  return r;
}

// warp-uniform span of 16-column chunks that can hold live keys for the 32 rows of lane quadrant `quad`
__device__ __forceinline__ void chunk_span(const PackParams& p, int quad, int& c_lo, int& c_hi) {
  const int nchunk = p.N >> 4;
  if (p.cross) { c_lo = 0; c_hi = nchunk; return; }
  const int g0 = (quad * 32) / p.Lq8;
  int g1 = (quad * 32 + 31) / p.Lq8;
  if (g1 > p.G - 1) g1 = p.G - 1;
  if (g0 > g1) { c_lo = c_hi = 0; return; }
  c_lo = (g0 * p.Lk8) >> 4;
  c_hi = min(nchunk, (g1 * p.Lk8 + p.Lk + 15) >> 4);
}

// ---------------------------------------------------------------------------------------------
// forward: grid (work items, H); 256 threads; TMEM = S (N columns), O aliases its first 64 columns
// ---------------------------------------------------------------------------------------------
template <int NPART>
__global__ void __launch_bounds__(NPART * 128, 2)
attn_pack_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const PackParams p) {
  const int item = blockIdx.x, h = blockIdx.y;
  if (p.cross && item >= __ldg(p.table)) return;  // CTA-uniform: the grid is sized for the worst case
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + p.off_k;
  uint8_t* sV = smem + p.off_v;
  const uint32_t sP_addr = sbase;  // P overwrites Q/K once S is in TMEM
  float* s_red = reinterpret_cast<float*>(smem + p.off_red);  // [NPART column parts][128 rows]
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_mma = bar_qk + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_qk + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, part = warp >> 2;  // TMEM lane quadrant (hardware rule) / column part of the row
  const int row = quad * 32 + lane;
  const int N = p.N, nchunk = N >> 4;
  const int q_rows = p.G * p.Lq8, k_rows = p.cross ? N : p.G * p.Lk8;
  constexpr int NT = NPART * 128;

  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    const int src = p.cross ? __ldg(p.table + p.items_off + item * PK_ITEM_INTS) : 0;
    int bs[PK_MAX_G];
#pragma unroll
    for (int g = 0; g < PK_MAX_G; ++g) bs[g] = item_seq(p, item, g);  // all table reads in flight before the first TMA
    mbar_arrive_expect_tx(bar_qk, (q_rows + k_rows) * 128);
    mbar_arrive_expect_tx(bar_v, k_rows * 128);
#pragma unroll
    for (int g = 0; g < PK_MAX_G; ++g) {
      if (g >= p.G) break;
      const int b = bs[g];
      tma_load_2d(sQ + g * p.Lq8 * 128, &tmap_q, bar_qk, h * 64, b >= 0 ? b * p.Lq : p.B * p.Lq);
      if (!p.cross) {
        const int kr = b >= 0 ? b * p.Lk : p.B * p.Lk;
        tma_load_2d(sK + g * p.Lk8 * 128, &tmap_k, bar_qk, h * 64, kr);
        tma_load_2d(sV + g * p.Lk8 * 128, &tmap_v, bar_v, h * 64, kr);
      }
    }
    if (p.cross) {
      tma_load_2d(sK, &tmap_k, bar_qk, h * 64, src * p.Lk);
      tma_load_2d(sV, &tmap_v, bar_v, h * 64, src * p.Lk);
    }
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  // rows no TMA box covers must not hold NaN bit patterns: they are multiplied by P = 0 / are tile padding
  zero_smem(sbase + q_rows * 128, (128 - q_rows) * 128, threadIdx.x, NT);
  if (!p.cross) {
    zero_smem(sbase + p.off_k + k_rows * 128, (N - k_rows) * 128, threadIdx.x, NT);
    zero_smem(sbase + p.off_v + k_rows * 128, (N - k_rows) * 128, threadIdx.x, NT);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (threadIdx.x == 0) {
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem, make_smem_desc(aq + k * 32, 16, 1024), make_smem_desc(ak + k * 32, 16, 1024), idesc, k != 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  const RowInfo ri = row_info(p, item, row, h);
  int c_lo, c_hi;
  chunk_span(p, quad, c_lo, c_hi);
  const bool warp_live = __any_sync(0xffffffffu, ri.ok) && c_hi > c_lo;
  const int cb = c_lo + ((c_hi - c_lo) * part) / NPART, ce = c_lo + ((c_hi - c_lo) * (part + 1)) / NPART;  // my live chunks
  mbar_wait_warp(bar_mma, 0);
  tc_fence_after();

  const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  float mx = -INFINITY, sum = 0.f;
  if (warp_live) {
    // pass 1: t = scale·qk + mask in the log2 domain, row max; t goes back to TMEM
    for (int c = cb; c < ce; ++c) {
      uint32_t s[16];
      tmem_ld_32x16(trow + c * 16, s);
      float add[16];
      const int kk0 = c * 16 - ri.klo;
      load_mask16(ri.mask_row, kk0, p.Lk, add);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int kk = kk0 + j;
        const float t = (ri.ok && kk >= 0 && kk < p.Lk) ? fmaf(__uint_as_float(s[j]), p.scale_log2, add[j]) : -INFINITY;
        mx = fmaxf(mx, t);
        s[j] = __float_as_uint(t);
      }
      tmem_st_32x16(trow + c * 16, s);
    }
    tmem_wait_st();
    s_red[part * 128 + row] = mx;
  }
  __syncthreads();  // max exchange; every S column has been read, so P may overwrite Q/K
  if (warp_live) {
#pragma unroll
    for (int i = 0; i < NPART; ++i) mx = fmaxf(mx, s_red[i * 128 + row]);
    if (mx == -INFINITY) mx = 0.f;
    const DropCfg dc = make_drop(p.dropout_p);
    const uint64_t doff = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
    for (int c = cb; c < ce; ++c) {
      uint32_t s[16];
      tmem_ld_32x16(trow + c * 16, s);
      tmem_wait_ld();
      float pr[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pr[j] = fast_exp2(__uint_as_float(s[j]) - mx);
        sum += pr[j];
      }
      if (p.dropout_p > 0.f) {
        const int kk0 = c * 16 - ri.klo;
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          const int kk = kk0 + j;
          if (ri.ok && kk >= 0 && kk < p.Lk) {
            float k[8];
            drop8(p.seed, doff, (ri.drop_base + kk) >> 3, dc, k);
#pragma unroll
            for (int i = 0; i < 8; ++i) pr[j + i] *= k[i];
          }
        }
      }
      st_shared_v4(sP_addr + swz_off(row, c * 16), pack_bf16x2(pr[0], pr[1]), pack_bf16x2(pr[2], pr[3]),
                   pack_bf16x2(pr[4], pr[5]), pack_bf16x2(pr[6], pr[7]));
      st_shared_v4(sP_addr + swz_off(row, c * 16 + 8), pack_bf16x2(pr[8], pr[9]), pack_bf16x2(pr[10], pr[11]),
                   pack_bf16x2(pr[12], pr[13]), pack_bf16x2(pr[14], pr[15]));
    }
  }
  // columns outside the live span (other sequences' keys) and rows of dead warps: P = 0
  for (int c = part; c < nchunk; c += NPART) {
    if (warp_live && c >= c_lo && c < c_hi) continue;
    st_shared_v4(sP_addr + swz_off(row, c * 16), 0u, 0u, 0u, 0u);
    st_shared_v4(sP_addr + swz_off(row, c * 16 + 8), 0u, 0u, 0u, 0u);
  }
  __syncthreads();  // every thread has read its partners' max before the slots are reused for the sums
  if (warp_live) s_red[part * 128 + row] = sum;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // O = P · V (accumulator aliases the first 64 S columns: all S reads are complete)
  if (threadIdx.x == 0) {
    mbar_wait(bar_v, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t av = smem_u32(sV);
    for (int ks = 0; ks < nchunk; ++ks)
      umma_bf16(tmem, make_smem_desc(sP_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                make_smem_desc(av + ks * 2048, 8192, 1024), idesc, ks != 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait_warp(bar_mma, 1);
  tc_fence_after();
  if (warp_live) {  // each part writes 64 / NPART of the 64 output dims of its rows
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NPART; ++i) tot += s_red[i * 128 + row];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;
    __nv_bfloat16* dst = p.o + (static_cast<int64_t>(ri.b) * p.Lq + ri.q) * p.ld_o + h * 64;
    store_row_part<NPART>(dst, trow, part, inv, ri.ok);
    if (ri.ok && part == 0) p.lse[(static_cast<int64_t>(ri.b) * p.H + h) * p.Lq + ri.q] = mx + log2f(tot);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// backward: grid (key groups, H) — self: one work item; cross: one K/V source with all its items.
// TMEM (512 columns): X = [0,256) holds S and dP (side by side when N <= 128, one after the other when
// N > 128) and afterwards dQ; dK tiles at 256 + 64t, dV tiles at 384 + 64t (t = key tile of 128).
// ---------------------------------------------------------------------------------------------

template <int NPART>
__global__ void __launch_bounds__(NPART * 128, 1)
attn_pack_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dk,
                     const __grid_constant__ CUtensorMap tmap_dv, const PackParams p) {
  const int grp = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // MMA issue runs warp-converged in warp 0 (index through a shuffle: provably uniform) with the instruction under
  // elect.sync, so that descriptors live in uniform registers (tools/mma_bench.cu: 58-67 instead of 91 cycles per MMA)
  const bool issuer_warp = __shfl_sync(0xffffffffu, warp, 0) == 0;
  const int quad = warp & 3, part = warp >> 2;
  const int row = quad * 32 + lane;
  constexpr int NT = NPART * 128;
  const int N = p.N, nchunk = N >> 4, ntile = (N + 127) >> 7;
  int first_item = grp, n_items = 1;
  if (p.cross) {
    first_item = __ldg(p.table + PK_HDR_INTS + grp);
    n_items = __ldg(p.table + PK_HDR_INTS + grp + 1) - first_item;
    if (n_items <= 0) {  // a K/V source nobody looked at: its gradient is zero (CTA-uniform exit)
      for (int i = threadIdx.x; i < p.Lk * 8; i += NT) {
        const int64_t r = static_cast<int64_t>(grp) * p.Lk + (i >> 3);
        const int c = (i & 7) * 8;
        *reinterpret_cast<uint4*>(p.dk + r * p.ld_dk + h * 64 + c) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(p.dv + r * p.ld_dv + h * 64 + c) = make_uint4(0u, 0u, 0u, 0u);
      }
      return;
    }
  }
  PK_TRACE(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* bar_kv = bar_q + 1;
  uint64_t* bar_mma = bar_q + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_q + 3);
  const int q_rows = p.G * p.Lq8, k_rows = p.cross ? N : p.G * p.Lk8;
  const bool dual = N <= 128;  // S and dP fit side by side

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_kv, 2 * k_rows * 128);
    if (p.cross) {
      tma_load_2d(smem + p.off_k, &tmap_k, bar_kv, h * 64, grp * p.Lk);
      tma_load_2d(smem + p.off_v, &tmap_v, bar_kv, h * 64, grp * p.Lk);
    } else {
      for (int g = 0; g < p.G; ++g) {
        const int b = item_seq(p, first_item, g);
        const int kr = b >= 0 ? b * p.Lk : p.B * p.Lk;
        tma_load_2d(smem + p.off_k + g * p.Lk8 * 128, &tmap_k, bar_kv, h * 64, kr);
        tma_load_2d(smem + p.off_v + g * p.Lk8 * 128, &tmap_v, bar_kv, h * 64, kr);
      }
    }
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  zero_smem(sbase + q_rows * 128, (128 - q_rows) * 128, threadIdx.x, NT);
  zero_smem(sbase + p.off_do + q_rows * 128, (128 - q_rows) * 128, threadIdx.x, NT);
  if (!p.cross) {
    zero_smem(sbase + p.off_k + k_rows * 128, (N - k_rows) * 128, threadIdx.x, NT);
    zero_smem(sbase + p.off_v + k_rows * 128, (N - k_rows) * 128, threadIdx.x, NT);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  PK_TRACE(1);
  const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const uint32_t aq = sbase, ado = sbase + p.off_do, ak = sbase + p.off_k, av = sbase + p.off_v;
  const uint32_t ap = sbase + p.off_p, ads = sbase + p.off_ds;
  const DropCfg dc = make_drop(p.dropout_p);
  const uint64_t doff = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
  const uint32_t idesc_s = make_idesc_bf16(128, N, 0, 0);
  const uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);
  const uint32_t idesc_dkv = make_idesc_bf16(128, 64, 1, 1);
  int c_lo, c_hi;
  chunk_span(p, quad, c_lo, c_hi);
  const int cb = c_lo + ((c_hi - c_lo) * part) / NPART, ce = c_lo + ((c_hi - c_lo) * (part + 1)) / NPART;
  uint32_t mma_phase = 0;

  for (int ci = 0; ci < n_items; ++ci) {
    const int item = first_item + ci;
    PK_TRACE(2 + ci * 10);
    if (threadIdx.x == 0) {
      int bs[PK_MAX_G];
#pragma unroll
      for (int g = 0; g < PK_MAX_G; ++g) bs[g] = item_seq(p, item, g);
      if (ci > 0) tma_store_wait_read();  // the previous item's dQ store has read its staging tile (= the Q tile)
      mbar_arrive_expect_tx(bar_q, 2 * q_rows * 128);
#pragma unroll
      for (int g = 0; g < PK_MAX_G; ++g) {
        if (g >= p.G) break;
        const int b = bs[g];
        const int qr = b >= 0 ? b * p.Lq : p.B * p.Lq;
        tma_load_2d(smem + g * p.Lq8 * 128, &tmap_q, bar_q, h * 64, qr);
        tma_load_2d(smem + p.off_do + g * p.Lq8 * 128, &tmap_do, bar_q, h * 64, qr);
      }
    }
    // per-row statistics while the tiles fly in: lse (log2 domain) and delta = rowsum(dO ∘ O)
    const RowInfo ri = row_info(p, item, row, h);
    const bool warp_live = __any_sync(0xffffffffu, ri.ok) && c_hi > c_lo;
    float my_lse = 0.f, my_delta = 0.f;
    if (ri.ok) {
      my_lse = p.lse[(static_cast<int64_t>(ri.b) * p.H + h) * p.Lq + ri.q];
    }
    if (ri.ok && p.delta) {
      my_delta = __ldg(p.delta + (static_cast<int64_t>(ri.b) * p.H + h) * p.Lq + ri.q);
    } else if (ri.ok) {
      const uint4* po = reinterpret_cast<const uint4*>(p.o + (static_cast<int64_t>(ri.b) * p.Lq + ri.q) * p.ld_o + h * 64);
      const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + (static_cast<int64_t>(ri.b) * p.Lq + ri.q) * p.ld_do + h * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 a = __ldg(po + i), d = __ldg(pd + i);
        my_delta += bf16_lo(a.x) * bf16_lo(d.x) + bf16_hi(a.x) * bf16_hi(d.x) + bf16_lo(a.y) * bf16_lo(d.y) +
                    bf16_hi(a.y) * bf16_hi(d.y) + bf16_lo(a.z) * bf16_lo(d.z) + bf16_hi(a.z) * bf16_hi(d.z) +
                    bf16_lo(a.w) * bf16_lo(d.w) + bf16_hi(a.w) * bf16_hi(d.w);
      }
    }
    PK_TRACE(3 + ci * 10);
    if (issuer_warp) {
      if (ci == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, ci & 1);
      tc_fence_after();
      PK_TRACE(4 + ci * 10);
      const uint64_t dq_ = make_smem_desc(aq, 16, 1024), dk_ = make_smem_desc(ak, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (elect_one()) umma_bf16(tmem, dq_ + 2 * k, dk_ + 2 * k, idesc_s, k != 0);
      if (dual) {
        const uint64_t ddo = make_smem_desc(ado, 16, 1024), dv_ = make_smem_desc(av, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (elect_one()) umma_bf16(tmem + 128, ddo + 2 * k, dv_ + 2 * k, idesc_s, k != 0);
      }
      if (elect_one()) umma_commit(bar_mma);
      __syncwarp();
    }
    mbar_wait_warp(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    PK_TRACE(5 + ci * 10);

    // ---- pass 1: P (and, when dP is already there, dS) for this thread's row and column half ----
    if (warp_live) {
#pragma unroll 1
      for (int c = cb; c < ce; ++c) {
        uint32_t s[16], dp[16];
        tmem_ld_32x16(trow + c * 16, s);
        if (dual) tmem_ld_32x16(trow + 128 + c * 16, dp);
        float add[16];
        const int kk0 = c * 16 - ri.klo;
        load_mask16(ri.mask_row, kk0, p.Lk, add);
        tmem_wait_ld();
        float pu[16], pd[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int kk = kk0 + j;
          const float t = fmaf(__uint_as_float(s[j]), p.scale_log2, add[j]) - my_lse;
          pu[j] = (ri.ok && kk >= 0 && kk < p.Lk) ? fast_exp2(t) : 0.f;
          pd[j] = pu[j];
        }
        float keep[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) keep[j] = 1.f;
        if (p.dropout_p > 0.f) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            const int kk = kk0 + j;
            if (ri.ok && kk >= 0 && kk < p.Lk) {
              float k[8];
              drop8(p.seed, doff, (ri.drop_base + kk) >> 3, dc, k);
#pragma unroll
              for (int i = 0; i < 8; ++i) { keep[j + i] = k[i]; pd[j + i] *= k[i]; }
            }
          }
        }
        float second[16];  // dual: dS; otherwise the un-dropped P parked in the dS tile until dP exists
#pragma unroll
        for (int j = 0; j < 16; ++j)
          second[j] = dual ? pu[j] * (__uint_as_float(dp[j]) * keep[j] - my_delta) : pu[j];
        const uint32_t o0 = swz_off(row, c * 16), o1 = swz_off(row, c * 16 + 8);
        st_shared_v4(ap + o0, pack_bf16x2(pd[0], pd[1]), pack_bf16x2(pd[2], pd[3]), pack_bf16x2(pd[4], pd[5]),
                     pack_bf16x2(pd[6], pd[7]));
        st_shared_v4(ap + o1, pack_bf16x2(pd[8], pd[9]), pack_bf16x2(pd[10], pd[11]), pack_bf16x2(pd[12], pd[13]),
                     pack_bf16x2(pd[14], pd[15]));
        st_shared_v4(ads + o0, pack_bf16x2(second[0], second[1]), pack_bf16x2(second[2], second[3]),
                     pack_bf16x2(second[4], second[5]), pack_bf16x2(second[6], second[7]));
        st_shared_v4(ads + o1, pack_bf16x2(second[8], second[9]), pack_bf16x2(second[10], second[11]),
                     pack_bf16x2(second[12], second[13]), pack_bf16x2(second[14], second[15]));
      }
    }
    // columns outside the live span / rows of dead warps contribute nothing: P = dS = 0
    for (int c = part; c < nchunk; c += NPART) {
      if (warp_live && c >= c_lo && c < c_hi) continue;
      const uint32_t o0 = swz_off(row, c * 16), o1 = swz_off(row, c * 16 + 8);
      st_shared_v4(ap + o0, 0u, 0u, 0u, 0u);
      st_shared_v4(ap + o1, 0u, 0u, 0u, 0u);
      st_shared_v4(ads + o0, 0u, 0u, 0u, 0u);
      st_shared_v4(ads + o1, 0u, 0u, 0u, 0u);
    }
    PK_TRACE(6 + ci * 10);
    if (!dual) {
      // ---- dP = dO · Vᵀ into the columns S occupied, then pass 2: dS = P ∘ (dP·keep − delta) ----
      tc_fence_before();
      __syncthreads();
      if (issuer_warp) {
        tc_fence_after();
        const uint64_t ddo = make_smem_desc(ado, 16, 1024), dv_ = make_smem_desc(av, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (elect_one()) umma_bf16(tmem, ddo + 2 * k, dv_ + 2 * k, idesc_s, k != 0);
        if (elect_one()) umma_commit(bar_mma);
        __syncwarp();
      }
      mbar_wait_warp(bar_mma, mma_phase);
      mma_phase ^= 1;
      tc_fence_after();
      PK_TRACE(7 + ci * 10);
      if (warp_live) {
#pragma unroll 1
        for (int c = cb; c < ce; ++c) {
          uint32_t dp[16];
          tmem_ld_32x16(trow + c * 16, dp);
          const uint32_t o0 = swz_off(row, c * 16), o1 = swz_off(row, c * 16 + 8);
          const uint4 u0 = ld_shared_v4(ads + o0), u1 = ld_shared_v4(ads + o1);
          const uint4 d0 = ld_shared_v4(ap + o0), d1 = ld_shared_v4(ap + o1);
          const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
          const uint32_t dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
          tmem_wait_ld();
          uint32_t out[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // the keep factor is recovered from the dropped copy: P_dropped != 0  <=>  kept (whenever P != 0)
            const float p0 = bf16_lo(uu[j]), p1 = bf16_hi(uu[j]);
            const float k0 = (dd[j] & 0xFFFFu) != 0u ? dc.scale : 0.f;
            const float k1 = (dd[j] & 0xFFFF0000u) != 0u ? dc.scale : 0.f;
            out[j] = pack_bf16x2(p0 * (__uint_as_float(dp[2 * j]) * k0 - my_delta),
                                 p1 * (__uint_as_float(dp[2 * j + 1]) * k1 - my_delta));
          }
          st_shared_v4(ads + o0, out[0], out[1], out[2], out[3]);
          st_shared_v4(ads + o1, out[4], out[5], out[6], out[7]);
        }
      }
    }
    PK_TRACE(8 + ci * 10);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    PK_TRACE(9 + ci * 10);

    if (issuer_warp) {
      tc_fence_after();
      // dQ = dS · K      (A: dS K-major over keys; B: K tile MN-major, N = 64 dims) -> X[0,64)
      {
        const uint64_t db0 = make_smem_desc(ak, 8192, 1024);
        for (int ks = 0; ks < nchunk; ++ks) {
          const uint64_t da = make_smem_desc(ads + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
          if (elect_one()) umma_bf16(tmem, da, db0 + ks * 128, idesc_dq, ks != 0);
        }
      }
      // dK[t] += dSᵀ · Q, dV[t] += Pᵀ · dO   (A: dS / P MN-major over keys, K = query rows; B: Q / dO MN-major)
      const int nqc = (q_rows + 15) >> 4;
      for (int t = 0; t < ntile; ++t) {
        const uint64_t das = make_smem_desc(ads + t * 32768, 16384, 1024), dbq = make_smem_desc(aq, 8192, 1024);
        for (int ks = 0; ks < nqc; ++ks)
          if (elect_one()) umma_bf16(tmem + p.tm_dk + t * 64, das + ks * 128, dbq + ks * 128, idesc_dkv, (ci | ks) != 0);
        const uint64_t dap = make_smem_desc(ap + t * 32768, 16384, 1024), dbo = make_smem_desc(ado, 8192, 1024);
        for (int ks = 0; ks < nqc; ++ks)
          if (elect_one()) umma_bf16(tmem + p.tm_dv + t * 64, dap + ks * 128, dbo + ks * 128, idesc_dkv, (ci | ks) != 0);
      }
      if (elect_one()) umma_commit(bar_mma);
      __syncwarp();
    }
    mbar_wait_warp(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    PK_TRACE(10 + ci * 10);
    // dQ rows -> bf16 staging in the (now idle) Q tile -> one TMA store per sequence slot (rows past Lq are clipped)
    if (warp_live) stage_row_part<NPART>(aq, row, trow, part, p.scale, row < q_rows);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();  // X, dO, P, dS are free for the next item; the Q tile once the bulk store has read it
    tc_fence_after();
    if (threadIdx.x == 0) {
      for (int g = 0; g < p.G; ++g) {
        const int bq = item_seq(p, item, g);
        if (bq >= 0) tma_store_3d(&tmap_dq, aq + g * p.Lq8 * 128, h * 64, 0, bq);
      }
      tma_store_commit();
    }
    PK_TRACE(11 + ci * 10);
  }

  // ---- drain dK / dV: thread == key row of tile t, each part stores 64 / NPART of the 64 dims ----
  // thread == key row of tile t: bf16 rows staged in the (dead) K / V tiles, written by TMA stores — cross: the source's
  // whole [N x 64] tile (rows past Lk clipped); self: one [Lk8 x 64] box per sequence slot
  for (int t = 0; t < ntile; ++t) {
    const int key = t * 128 + row;  // row of the K / V tile in shared memory (rows are 128 B apart in both modes)
    stage_row_part<NPART>(ak + t * 16384, row, trow + p.tm_dk + t * 64, part, p.scale, key < N);
    stage_row_part<NPART>(av + t * 16384, row, trow + p.tm_dv + t * 64, part, 1.0f, key < N);
  }
  PK_TRACE(62);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p.cross) {
      tma_store_3d(&tmap_dk, ak, h * 64, 0, grp);
      tma_store_3d(&tmap_dv, av, h * 64, 0, grp);
    } else {
      for (int g = 0; g < p.G; ++g) {
        const int bk = item_seq(p, first_item, g);
        if (bk >= 0) {
          tma_store_3d(&tmap_dk, ak + g * p.Lk8 * 128, h * 64, 0, bk);
          tma_store_3d(&tmap_dv, av + g * p.Lk8 * 128, h * 64, 0, bk);
        }
      }
    }
    tma_store_commit();
    tma_store_wait_read();  // shared memory must stay intact until the bulk stores (dQ of the last item too) have read it
  }
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
  PK_TRACE(63);
}

// ---------------------------------------------------------------------------------------------
// work-item table for cross mode: group the B query sequences by K/V source, G per item.
//   table = {n_items, G, n_kv, B} | first[n_kv + 1] (items of source s: [first[s], first[s+1])) | pad to 4 |
//           items[n_items_max][12] = {src, n, b[0..7], 0, 0}
// One CTA; ranks are computed by counting (deterministic: a sequence's slot never depends on thread timing).
// ---------------------------------------------------------------------------------------------
constexpr int GB_THREADS = 1024;
constexpr int GB_MAX_KV = 4096;

__global__ void __launch_bounds__(GB_THREADS, 1)
attn_group_build_kernel(const int32_t* __restrict__ kv_index, int B, int n_kv, int G, int n_items_max, int32_t* __restrict__ table) {
  __shared__ int cnt[GB_MAX_KV];
  __shared__ int first[GB_MAX_KV + 1];
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x;
  int32_t* items = table + PK_HDR_INTS + ((n_kv + 1 + 3) & ~3);
  for (int s = tid; s < n_kv; s += GB_THREADS) cnt[s] = 0;
  for (int i = tid; i < n_items_max * PK_ITEM_INTS; i += GB_THREADS) {
    const int f = i % PK_ITEM_INTS;
    items[i] = (f >= 2 && f < 2 + PK_MAX_G) ? -1 : 0;
  }
  __syncthreads();
  for (int b = tid; b < B; b += GB_THREADS) {
    const int s = kv_index[b];
    if (s >= 0 && s < n_kv) atomicAdd(&cnt[s], 1);
  }
  __syncthreads();
  // exclusive scan of ceil(cnt / G) over the sources: contiguous span per thread, then a block scan of the spans
  const int span = (n_kv + GB_THREADS - 1) / GB_THREADS;
  const int s0 = tid * span, s1 = min(n_kv, s0 + span);
  int local = 0;
  for (int s = s0; s < s1; ++s) local += (cnt[s] + G - 1) / G;
  int incl = local;
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += v;
  }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = warp_tot[tid];
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, d);
      if (tid >= d) w += v;
    }
    warp_tot[tid] = w;  // inclusive totals of the warps
  }
  __syncthreads();
  int run = incl - local + ((tid >> 5) ? warp_tot[(tid >> 5) - 1] : 0);
  for (int s = s0; s < s1; ++s) {
    first[s] = run;
    run += (cnt[s] + G - 1) / G;
  }
  if (tid == GB_THREADS - 1) first[n_kv] = warp_tot[31];
  __syncthreads();
  for (int s = tid; s <= n_kv; s += GB_THREADS) table[PK_HDR_INTS + s] = first[s];
  if (tid == 0) {
    table[0] = first[n_kv];
    table[1] = G;
    table[2] = n_kv;
    table[3] = B;
  }
  for (int b = tid; b < B; b += GB_THREADS) {
    const int s = kv_index[b];
    if (s < 0 || s >= n_kv) continue;
    int rank = 0;
    for (int j = 0; j < b; ++j) rank += (kv_index[j] == s) ? 1 : 0;
    const int it = first[s] + rank / G, slot = rank % G;
    int32_t* rec = items + it * PK_ITEM_INTS;
    rec[2 + slot] = b;
    if (slot == 0) {
      rec[0] = s;
      rec[1] = min(G, cnt[s] - (rank / G) * G);
    }
  }
}

// threads per CTA / 128 = how many warps share a TMEM lane quadrant (each takes a column part of the rows).
// Measured on B200 (tools/profile_attn.py, round 2, after the drains became TMA stores): the FORWARD is faster with 2 parts
// (57.9 / 90.5 / 171.5 us against 65.2 / 109.9 / 172.2 us with 4 for text B=256, fusion self B=576, grouped cross B=576: two
// 256-thread CTAs share an SM), the BACKWARD with 4 (117.6 / 214.3 / 343.7 us against 126.0 / 228.5 / 395.3 us: one CTA per SM,
// the softmax / dS passes are its longest phases and split four ways).
// Tuning override: X2K_PACK_PARTS_FWD / X2K_PACK_PARTS_BWD = 2 | 4 (read once).
int pack_parts(bool backward) {
  static int cached[2] = {0, 0};
  int& c = cached[backward ? 1 : 0];
  if (c == 0) {
    const char* e = getenv(backward ? "X2K_PACK_PARTS_BWD" : "X2K_PACK_PARTS_FWD");
    const int dflt = backward ? 4 : 2;
    c = (e && (atoi(e) == 2 || atoi(e) == 4)) ? atoi(e) : dflt;
  }
  return c;
}

int pack_slots(int Lq) {
  const int Lq8 = (Lq + 7) & ~7;
  if (Lq <= 0 || Lq8 > 64) return 0;
  return min(128 / Lq8, PK_MAX_G);
}

// Decide whether x2k_attn_fwd / x2k_attn_bwd should take the packed path and fill the launch geometry.
// Returns 1 = packed, 0 = use attn.cu's kernels, < 0 = error.
int attn_pack_plan(const X2kAttnArgs& a, bool backward, PackParams& p) {
  if (a.bias != nullptr || (backward && a.ds_out != nullptr)) return 0;
  int G = pack_slots(a.Lq);
  if (G < 1) return 0;
  const int Lq8 = (a.Lq + 7) & ~7, Lk8 = (a.Lk + 7) & ~7;
  const bool cross = a.kv_groups != nullptr;
  if (!cross) {
    if (a.kv_index != nullptr) return 0;
    if (a.n_kv != 0 && a.n_kv != a.B) return 0;
    G = min(G, 256 / Lk8);
    if (G < 2) return 0;  // nothing to pack
  } else {
    X2K_REQUIRE(a.n_kv > 0, "x2k_attn: kv_groups needs n_kv");
    if (((a.Lk + 15) & ~15) > 256) return 0;
  }
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.Lq8 = Lq8; p.Lk8 = Lk8; p.G = G;
  p.N = cross ? ((a.Lk + 15) & ~15) : ((G * Lk8 + 15) & ~15);
  p.Lk_pad = (a.Lk + 15) & ~15;
  p.cross = cross ? 1 : 0;
  p.n_kv = cross ? a.n_kv : a.B;
  p.table = a.kv_groups;
  p.items_off = PK_HDR_INTS + ((p.n_kv + 1 + 3) & ~3);
  p.scale = a.scale; p.scale_log2 = a.scale * kLog2e;
  p.mask = a.mask; p.mask_b_stride = a.mask_b_stride; p.mask_q_stride = a.mask_q_stride;
  p.dropout_p = a.dropout_p; p.seed = a.dropout_seed; p.offset = a.dropout_offset; p.offset_dev = a.dropout_offset_dev;
  p.o = static_cast<__nv_bfloat16*>(a.o); p.ld_o = a.ld_o; p.lse = a.lse;
  p.d_o = static_cast<const __nv_bfloat16*>(a.d_o); p.ld_do = a.ld_do;
  p.delta = backward ? a.delta_ws : nullptr;
  p.dq = static_cast<__nv_bfloat16*>(a.dq); p.dk = static_cast<__nv_bfloat16*>(a.dk); p.dv = static_cast<__nv_bfloat16*>(a.dv);
  p.ld_dq = a.ld_dq; p.ld_dk = a.ld_dk; p.ld_dv = a.ld_dv;
  const uint32_t kv_bytes = (static_cast<uint32_t>(p.N) * 128 + 1023) & ~1023u;
  // backward: dK/dV read P/dS as whole 128-key tiles (2 blocks of 64 columns each)
  const uint32_t p_bytes = backward ? static_cast<uint32_t>((p.N + 127) >> 7) * 32768 : static_cast<uint32_t>((p.N + 63) >> 6) * 16384;
  if (!backward) {
    p.off_do = 0;
    p.off_k = 16384;
    p.off_v = max(16384 + kv_bytes, p_bytes);
    p.off_p = 0; p.off_ds = 0;
    p.off_red = p.off_v + kv_bytes;
    p.off_bar = p.off_red + 2048;
    p.tmem_cols = p.N <= 128 ? 128 : 256;
  } else {
    p.off_do = 16384;
    p.off_k = 32768;
    p.off_v = p.off_k + kv_bytes;
    p.off_p = p.off_v + kv_bytes;
    p.off_ds = p.off_p + p_bytes;
    p.off_red = p.off_ds + p_bytes;
    p.off_bar = p.off_red;
    p.tmem_cols = 512;
    p.tm_dk = 256;
    p.tm_dv = 384;
  }
  return 1;
}

int attn_pack_fwd_launch(const X2kAttnArgs& a, PackParams& p, cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  int rc;
  const uint64_t kv_rows = static_cast<uint64_t>(p.n_kv) * a.Lk;
  const uint32_t kbox = p.cross ? p.N : p.Lk8;
  if ((rc = make_tmap_bf16_2d(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_q, p.Lq8, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, a.k, kv_rows, static_cast<uint64_t>(a.H) * 64, a.ld_k, kbox, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, a.v, kv_rows, static_cast<uint64_t>(a.H) * 64, a.ld_v, kbox, 64))) return rc;
  const int smem = p.off_bar + 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_pack_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_pack_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
    attr_set = true;
  }
  X2K_REQUIRE(smem <= 116 * 1024, "x2k_attn_fwd (packed): %d bytes of shared memory", smem);
  const int n_items = p.cross ? (a.B + p.n_kv * (p.G - 1)) / p.G : (a.B + p.G - 1) / p.G;
  dim3 grid(n_items, a.H);
  if (pack_parts(false) == 4)
    attn_pack_fwd_kernel<4><<<grid, 512, smem, stream>>>(tq, tk, tv, p);
  else
    attn_pack_fwd_kernel<2><<<grid, 256, smem, stream>>>(tq, tk, tv, p);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

int attn_pack_bwd_launch(const X2kAttnArgs& a, PackParams& p, cudaStream_t stream) {
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  const uint64_t kv_rows = static_cast<uint64_t>(p.n_kv) * a.Lk;
  const uint32_t kbox = p.cross ? p.N : p.Lk8;
  if ((rc = make_tmap_bf16_2d(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_q, p.Lq8, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tk, a.k, kv_rows, static_cast<uint64_t>(a.H) * 64, a.ld_k, kbox, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tv, a.v, kv_rows, static_cast<uint64_t>(a.H) * 64, a.ld_v, kbox, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, a.d_o, static_cast<uint64_t>(a.B) * a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_do, p.Lq8, 64))) return rc;
  // gradients leave as TMA stores of whole staging tiles; the per-sequence views clip the rows past Lq / Lk
  CUtensorMap tdq, tdk, tdv;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  X2K_REQUIRE(al16(a.dq) && al16(a.dk) && al16(a.dv), "x2k_attn_bwd (packed): dq / dk / dv must be 16-byte aligned");
  if ((rc = make_tmap_bf16_seq3d(&tdq, a.dq, a.B, a.Lq, static_cast<uint64_t>(a.H) * 64, a.ld_dq, p.Lq8))) return rc;
  if ((rc = make_tmap_bf16_seq3d(&tdk, a.dk, p.n_kv, a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_dk, kbox))) return rc;
  if ((rc = make_tmap_bf16_seq3d(&tdv, a.dv, p.n_kv, a.Lk, static_cast<uint64_t>(a.H) * 64, a.ld_dv, kbox))) return rc;
  const int smem = p.off_bar + 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_pack_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    X2K_CHECK_CUDA(cudaFuncSetAttribute(attn_pack_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  X2K_REQUIRE(smem <= 227 * 1024, "x2k_attn_bwd (packed): %d bytes of shared memory", smem);
  dim3 grid(p.cross ? p.n_kv : (a.B + p.G - 1) / p.G, a.H);
  if (pack_parts(true) == 4)
    attn_pack_bwd_kernel<4><<<grid, 512, smem, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  else
    attn_pack_bwd_kernel<2><<<grid, 256, smem, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

}  // namespace

// entry points used by attn.cu's dispatch: 1 = not eligible (caller falls back to the one-sequence-per-tile kernels)
int attn_pack_fwd(const X2kAttnArgs& a, cudaStream_t stream) {
  PackParams p;
  const int plan = attn_pack_plan(a, false, p);
  if (plan <= 0) return plan < 0 ? plan : 1;
  return attn_pack_fwd_launch(a, p, stream);
}
int attn_pack_bwd(const X2kAttnArgs& a, cudaStream_t stream) {
  PackParams p;
  const int plan = attn_pack_plan(a, true, p);
  if (plan <= 0) return plan < 0 ? plan : 1;
  return attn_pack_bwd_launch(a, p, stream);
}

}  // namespace x2k

using namespace x2k;

#ifdef X2K_PACK_TRACE
extern "C" int x2k_debug_pack_trace(long long* out128) {
  return cudaMemcpyFromSymbol(out128, g_pack_trace, sizeof(long long) * 128) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int32_t x2k_attn_group_slots(int32_t Lq) { return pack_slots(Lq); }

extern "C" int64_t x2k_attn_group_table_ints(int32_t B, int32_t n_kv, int32_t Lq) {
  const int G = pack_slots(Lq);
  if (G < 1 || B <= 0 || n_kv <= 0) return 0;
  const int64_t n_items_max = (static_cast<int64_t>(B) + static_cast<int64_t>(n_kv) * (G - 1)) / G;
  return PK_HDR_INTS + ((n_kv + 1 + 3) & ~3) + n_items_max * PK_ITEM_INTS;
}

extern "C" int x2k_attn_group_build(const int32_t* kv_index, int32_t B, int32_t n_kv, int32_t Lq, int32_t* table,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(kv_index && table && B > 0 && n_kv > 0, "x2k_attn_group_build: bad arguments");
  const int G = pack_slots(Lq);
  X2K_REQUIRE(G >= 1, "x2k_attn_group_build: Lq=%d is not eligible for the packed path (Lq <= 64)", Lq);
  X2K_REQUIRE(n_kv <= GB_MAX_KV, "x2k_attn_group_build: n_kv=%d > %d", n_kv, GB_MAX_KV);
  const int n_items_max = (B + n_kv * (G - 1)) / G;
  attn_group_build_kernel<<<1, GB_THREADS, 0, stream>>>(kv_index, B, n_kv, G, n_items_max, table);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}
