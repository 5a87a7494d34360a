// gemm.cu — persistent, warp-specialised tcgen05 GEMM with fused epilogues for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] · B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// Pipeline (one CTA per SM, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled tiles into a STAGES-deep smem ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x4 per
//               k-block; tcgen05.commit releases smem slots and signals the epilogue
//   warps 2..9  epilogue: tcgen05.ld the fp32 accumulator (thread == output row), apply the fused
//               epilogue (bias / GELU / GELU' / dropout / LayerScale / DropPath / residual /
//               accumulate), vectorised global stores.  Two TMEM accumulator stages let the epilogue
//               of tile i overlap the MMAs of tile i+1.
// Operand majors: K-major (row-major [rows, K]) or MN-major (row-major [K, rows]) for A and B,
// selected through the UMMA instruction/smem descriptors — forward, dgrad and wgrad GEMMs all
// read the tensors where they lie, no transposes are materialised.
//
// Replaces the F.linear(+bias+act+residual) sequences cited in include/x2k.h.
//
// X2K_BUILD_PARTS: 4   (compiled 4x with -DX2K_GEMM_PART=0..3; each part instantiates a quarter of the epilogue
//                       variants so the ~130 kernel instantiations build in parallel)
#include <cstdlib>

#include "common.cuh"

namespace x2k {

struct EpiParams {
  int M, N, K;
  const float* bias;
  int act;
  const __nv_bfloat16* aux;
  int64_t ld_aux;
  __nv_bfloat16* preact_out;
  int64_t ld_preact;
  float dropout_p;
  uint64_t dropout_seed, dropout_offset;
  const uint64_t* dropout_offset_dev;  // added to dropout_offset when set (advanced on the device between graph replays)
  const float* gamma;
  const float* row_scale;
  int rows_per_scale;
  const float* residual;
  int64_t ld_res;
  int accumulate;
  int split_k;  // >1: work unit = (tile, k-slice), fp32 output accumulated with atomics
  int use_pair;  // host-side: launch the CTA-pair (cta_group::2) kernel
  int raster_m;  // 1: consecutive tiles walk down M (CTAs running together share the B tile), 0: walk along N (share A)
  __nv_bfloat16* out_bf16;
  int64_t ld_out_bf16;
  float* out_f32;
  int64_t ld_out_f32;
  // debug overrides for the MN-major smem descriptors (0 = defaults); env X2K_DBG_MN="lbo,sbo,kadv"
  uint32_t dbg_lbo, dbg_sbo, dbg_kadv;
  int tma_epi;  // host-side decision: residual tiles TMA-loaded / fp32 output tiles TMA-stored (pair kernel, see epilogue_warp_tma)
  // fused cross entropy over the N (vocabulary) dimension, generic epilogue only (X2kGemmArgs.ce_*)
  int ce_mode;                // 0 off, 1 softmax statistics (no logits are stored), 2 gradient (exp(v - lse) - onehot) * row_grad
  const int64_t* ce_labels;   // [M], negative = ignored row
  float* ce_partials;         // mode 1: [M, ceil(N/16), 2] (max, sum exp(v - max)) per 16-column group
  float* ce_target_logit;     // mode 1: [M] logit of the label
  const float* ce_lse;        // mode 2: [M] log-sum-exp from x2k_ce_finalize
  const float* ce_row_grad;   // mode 2: [M] upstream gradient of the per-row loss
};

// natural-log online-softmax statistics of 16 logits (columns >= n_valid excluded)
__device__ __forceinline__ void ce_stats16(const float (&v)[16], int n_valid, float& mx, float& sum) {
  mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < n_valid) mx = fmaxf(mx, v[j]);
  sum = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < n_valid) sum += fast_exp2((v[j] - mx) * 1.4426950408889634f);
}
// fused-CE work on the 16 columns [n0, n0+16) of row m (after the bias add); returns through v in mode 2
__device__ __forceinline__ void ce_apply16(const EpiParams& p, int m, int n0, float (&v)[16]) {
  if (m >= p.M) return;
  const int n_valid = min(16, p.N - n0);
  const long long lbl = p.ce_labels ? p.ce_labels[m] : -1;
  const int t = (lbl >= n0 && lbl < n0 + n_valid) ? static_cast<int>(lbl - n0) : -1;
  if (p.ce_mode == 1) {
    float mx, sum;
    ce_stats16(v, n_valid, mx, sum);
    const int64_t groups = (p.N + 15) >> 4;
    float2* dst = reinterpret_cast<float2*>(p.ce_partials) + static_cast<int64_t>(m) * groups + (n0 >> 4);
    *dst = make_float2(mx, sum);
    if (t >= 0) {
      float tl = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) tl = (j == t) ? v[j] : tl;
      p.ce_target_logit[m] = tl;
    }
  } else {
    const float lse = __ldg(p.ce_lse + m), g = lbl >= 0 ? __ldg(p.ce_row_grad + m) : 0.0f;  // ignored rows: loss == 0
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = g * (fast_exp2((v[j] - lse) * 1.4426950408889634f) - (j == t ? 1.0f : 0.0f));
  }
}

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16;  // 4 per TMEM lane quadrant, each takes a quarter of the tile's columns
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB

template <int BLOCK_N>
struct Cfg {
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : 6;  // 192 KB of operand stages + 32 KB of epilogue tiles
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages
  static constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 32 * 64;  // per-warp transpose tile (EPI_TILE_BYTES)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};


// Apply the fused epilogue to the 16 columns [n0, n0+16) of row m with thread-per-row global accesses and a bounds
// test per element: only the ragged last chunk of an N % 32 != 0 output (the vocabulary GEMM, N < 32 heads) comes here.
__device__ __forceinline__ void epilogue_tail16(const EpiParams& p, int m, int n0, uint32_t (&acc)[16]) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
  }
  if (p.ce_mode) ce_apply16(p, m, n0, v);
  if (p.act == X2K_ACT_GELU_SAVE_GRAD) {  // preact_out receives GELU'(v), v becomes GELU(v)
    __nv_bfloat16* dst = p.preact_out + static_cast<int64_t>(m) * p.ld_preact + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float g, dg;
      gelu_erf_both(v[j], g, dg);
      v[j] = g;
      if (n0 + j < p.N) dst[j] = __float2bfloat16_rn(dg);
    }
  } else if (p.preact_out) {
    __nv_bfloat16* dst = p.preact_out + static_cast<int64_t>(m) * p.ld_preact + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < p.N) dst[j] = __float2bfloat16_rn(v[j]);
  }
  if (p.act == X2K_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
  } else if (p.act == X2K_ACT_MUL_AUX || p.act == X2K_ACT_GELU_BWD) {
    const __nv_bfloat16* src = p.aux + static_cast<int64_t>(m) * p.ld_aux + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < p.N) {
        const float a = __bfloat162float(src[j]);
        v[j] *= p.act == X2K_ACT_MUL_AUX ? a : gelu_erf_grad(a);
      }
  }
  if (p.dropout_p > 0.0f) {
    const DropCfg dc = make_drop(p.dropout_p);
    const uint64_t base = static_cast<uint64_t>(m) * static_cast<uint64_t>(p.N) + static_cast<uint64_t>(n0);
    const uint64_t doff = p.dropout_offset + (p.dropout_offset_dev ? __ldg(p.dropout_offset_dev) : 0ull);
    // N is a multiple of 8 whenever dropout is used (checked on the host), so base % 8 == 0.
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      float k[8];
      drop8(p.dropout_seed, doff, (base + j) >> 3, dc, k);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[j + i] *= k[i];
    }
  }
  if (p.gamma) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < p.N) v[j] *= __ldg(p.gamma + n0 + j);
  }
  if (p.row_scale) {
    const float s = __ldg(p.row_scale + m / p.rows_per_scale);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= s;
  }
  if (p.residual) {
    const float* src = p.residual + static_cast<int64_t>(m) * p.ld_res + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < p.N) v[j] += __ldg(src + j);
  }
  if (p.out_f32) {
    float* dst = p.out_f32 + static_cast<int64_t>(m) * p.ld_out_f32 + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < p.N) {
        if (p.split_k > 1) atomicAdd(dst + j, v[j]);
        else dst[j] = p.accumulate ? dst[j] + v[j] : v[j];
      }
  }
  if (p.out_bf16) {
    __nv_bfloat16* dst = p.out_bf16 + static_cast<int64_t>(m) * p.ld_out_bf16 + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < p.N) dst[j] = __float2bfloat16_rn(v[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// Coalesced epilogue.  tcgen05.ld hands every thread one ROW of the accumulator, but a warp-wide global access in
// which each lane touches a different row costs 32 L1 wavefronts / 32 half-filled sectors per instruction.  Every
// per-element tensor of the epilogue (pre-activation, aux, residual, outputs) therefore goes through a per-warp
// transpose tile in shared memory: 32 rows x 64 payload bytes (32 bf16 or 16 fp32 columns), rows 64 bytes apart, the
// 16-byte chunk c of row r stored at chunk c ^ ((r >> 1) & 3).  That XOR keeps 16-byte accesses conflict-free on both
// sides: thread-per-row (8 lanes = 8 rows, one chunk each: even rows fall into one 64-byte bank half with four
// different physical chunks, odd rows into the other) and global side (lane l handles chunk l%4 of row 8*it + l/4:
// 8 lanes = 2 rows x 4 chunks).  On the global side every instruction moves 8 rows x 64 contiguous bytes.
// fp32 tensors are processed as 16-column halves.
// ---------------------------------------------------------------------------------------------
constexpr int EPI_TILE_BYTES = 32 * 64;  // 2 KB per epilogue warp

__device__ __forceinline__ uint32_t epi_off(int r, int c) { return static_cast<uint32_t>(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// thread-per-row side: this thread's 16 words <-> row `lane` of the tile
__device__ __forceinline__ void tile_put(uint32_t sa, int lane, const uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) sts128(sa + epi_off(lane, j), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  __syncwarp();
}
__device__ __forceinline__ void tile_get(uint32_t sa, int lane, uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 q = lds128(sa + epi_off(lane, j));
    w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w;
  }
  __syncwarp();
}
// global side: tile <-> 32 rows x 64 bytes at `g` (row pitch ld_bytes); rows >= `rows` are skipped / zero-filled
__device__ __forceinline__ void tile_store(uint32_t sa, int lane, uint8_t* g, int64_t ld_bytes, int rows, bool atomic_f32) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), c = lane & 3;
    const uint4 q = lds128(sa + epi_off(r, c));
    if (r < rows) {
      uint8_t* dst = g + r * ld_bytes + c * 16;
      if (atomic_f32)
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z),
                                                               __uint_as_float(q.w)));
      else
        *reinterpret_cast<uint4*>(dst) = q;
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void tile_load(uint32_t sa, int lane, const uint8_t* g, int64_t ld_bytes, int rows) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), c = lane & 3;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows) q = *reinterpret_cast<const uint4*>(g + r * ld_bytes + c * 16);
    sts128(sa + epi_off(r, c), q.x, q.y, q.z, q.w);
  }
  __syncwarp();
}
__device__ __forceinline__ void pack32(const float (&v)[32], uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
}

// Compile-time epilogue feature mask.  The kernel is instantiated for the handful of combinations the hot path
// uses (the epilogue warps are latency-bound: every branch and address computation of an unused feature costs issue
// slots on the critical path of a K = 768 tile) plus a generic variant that tests the runtime pointers.
enum : int {
  EF_BIAS = 1, EF_PREACT = 2, EF_GELU = 4, EF_GELU_BWD = 8, EF_DROPOUT = 16, EF_SCALE = 32 /*gamma and/or row_scale*/,
  EF_RESIDUAL = 64, EF_OUT_F32 = 128, EF_OUT_BF16 = 256, EF_GELU_SAVE = 512 /*GELU + stored GELU'*/, EF_MUL_AUX = 1024,
  EF_GENERIC = 1 << 20
};
template <int EPI, int F>
__device__ __forceinline__ bool has(bool runtime) {
  if constexpr (EPI == EF_GENERIC) return runtime;
  else return (EPI & F) != 0;
}

// Fused epilogue of a full 32-column chunk [n0, n0+32) for the 32 rows [mw, mw+32) owned by this warp (thread == row
// mw+lane; `rows` = number of those rows inside the matrix), read from TMEM at `taddr`.
//
// The CTA runs 16 epilogue warps (4 per scheduler): with 8, the K = 768 GEMMs of the step were bound by the latency of
// their epilogue chains (ncu, profiles/: issue slots 17-40% busy, long-scoreboard stalls on the residual loads, tensor
// pipe 18-36% busy while the plain bf16 epilogue reached 73%).  18 warps leave ~100 registers per thread, so the chunk
// is processed as two 16-column halves: fp32 tensors (residual, out_f32) move one half = one 64-byte tile row at a
// time anyway; bf16 tensors (pre-activation, aux, out_bf16) are packed per half and cross the transpose tile once per
// chunk (32 bf16 columns = 64 bytes per row).  Same arithmetic, in the same order, as epilogue_tail16.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk32(const EpiParams& p, uint32_t sa, int lane, int mw, int rows, int n0,
                                                 uint32_t taddr) {
  const int m = mw + lane;
  uint32_t wpre[16], wout[16], waux[16], w[16];
  constexpr bool kGeneric = EPI == EF_GENERIC;
  const bool f_save = has<EPI, EF_GELU_SAVE>(p.act == X2K_ACT_GELU_SAVE_GRAD);
  const bool f_pre = has<EPI, EF_PREACT>(p.preact_out != nullptr) || f_save;
  const bool f_gelu = has<EPI, EF_GELU>(p.act == X2K_ACT_GELU);
  const bool f_mul = has<EPI, EF_MUL_AUX>(p.act == X2K_ACT_MUL_AUX);
  const bool f_gbwd = has<EPI, EF_GELU_BWD>(p.act == X2K_ACT_GELU_BWD);
  (void)kGeneric;
  if (f_mul || f_gbwd) {
    tile_load(sa, lane, reinterpret_cast<const uint8_t*>(p.aux + static_cast<int64_t>(mw) * p.ld_aux + n0), p.ld_aux * 2, rows);
    tile_get(sa, lane, waux);
  }
  // light epilogues (no activation math, dropout or residual) have registers to spare: both halves are fetched from
  // TMEM up front, so the second load is in flight while the first half is processed
  constexpr bool kPrefetch = EPI != EF_GENERIC && (EPI & (EF_GELU | EF_GELU_SAVE | EF_GELU_BWD | EF_DROPOUT | EF_RESIDUAL)) == 0;
  uint32_t acc2[2][16];
  if (kPrefetch) {
    tmem_ld_32x16(taddr, acc2[0]);
    tmem_ld_32x16(taddr + 16, acc2[1]);
  }
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int nh = n0 + 16 * hf;
    float v[16];
    if (kPrefetch) {
      if (hf == 0) tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc2[hf][j]);
    } else {
      uint32_t acc[16];
      tmem_ld_32x16(taddr + 16 * hf, acc);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
    }
    if (has<EPI, EF_BIAS>(p.bias != nullptr)) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nh + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if constexpr (kGeneric) {
      if (p.ce_mode) ce_apply16(p, m, nh, v);
    }
    if (f_save) {
      // staged 8 elements at a time (all reciprocals / exponentials, then the polynomials): 8 independent
      // MUFU -> FMA chains in flight per thread
#pragma unroll
      for (int hb = 0; hb < 16; hb += 8) {
        float hh[8], ee[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) gelu_parts(v[hb + j], hh[j], ee[j]);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float x0 = v[hb + j], x1 = v[hb + j + 1];
          wpre[8 * hf + ((hb + j) >> 1)] =
              pack_bf16x2(gelu_grad_from_parts(x0, hh[j], ee[j]), gelu_grad_from_parts(x1, hh[j + 1], ee[j + 1]));
          v[hb + j] = fmaf(-fabsf(x0), hh[j], fmaxf(x0, 0.0f));
          v[hb + j + 1] = fmaf(-fabsf(x1), hh[j + 1], fmaxf(x1, 0.0f));
        }
      }
    } else if (f_pre) {
#pragma unroll
      for (int j = 0; j < 8; ++j) wpre[8 * hf + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    }
    if (f_gelu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
    } else if (f_mul) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[2 * j] *= bf16_lo(waux[8 * hf + j]);
        v[2 * j + 1] *= bf16_hi(waux[8 * hf + j]);
      }
    } else if (f_gbwd) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[2 * j] *= gelu_erf_grad(bf16_lo(waux[8 * hf + j]));
        v[2 * j + 1] *= gelu_erf_grad(bf16_hi(waux[8 * hf + j]));
      }
    }
    if (has<EPI, EF_DROPOUT>(true) && p.dropout_p > 0.0f) {
      const DropCfg dc = make_drop(p.dropout_p);
      const uint64_t base = static_cast<uint64_t>(m) * static_cast<uint64_t>(p.N) + static_cast<uint64_t>(nh);
      const uint64_t doff = p.dropout_offset + (p.dropout_offset_dev ? __ldg(p.dropout_offset_dev) : 0ull);
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        float k[8];
        drop8(p.dropout_seed, doff, (base + j) >> 3, dc, k);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[j + i] *= k[i];
      }
    }
    if (has<EPI, EF_SCALE>(true) && p.gamma) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + nh + j));
        v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
      }
    }
    if (has<EPI, EF_SCALE>(true) && p.row_scale) {
      const float sc = __ldg(p.row_scale + min(m, p.M - 1) / p.rows_per_scale);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= sc;
    }
    if (has<EPI, EF_RESIDUAL>(p.residual != nullptr)) {
      tile_load(sa, lane, reinterpret_cast<const uint8_t*>(p.residual + static_cast<int64_t>(mw) * p.ld_res + nh), p.ld_res * 4,
                rows);
      tile_get(sa, lane, w);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(w[j]);
    }
    if (has<EPI, EF_OUT_F32>(p.out_f32 != nullptr)) {
      const bool atomic = p.split_k > 1;
      uint8_t* dst = reinterpret_cast<uint8_t*>(p.out_f32 + static_cast<int64_t>(mw) * p.ld_out_f32 + nh);
      if (p.accumulate && !atomic) {
        tile_load(sa, lane, dst, p.ld_out_f32 * 4, rows);
        tile_get(sa, lane, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(w[j]);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(v[j]);
      tile_put(sa, lane, w);
      tile_store(sa, lane, dst, p.ld_out_f32 * 4, rows, atomic);
    }
    if (has<EPI, EF_OUT_BF16>(p.out_bf16 != nullptr)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) wout[8 * hf + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    }
  }
  if (f_pre) {
    tile_put(sa, lane, wpre);
    tile_store(sa, lane, reinterpret_cast<uint8_t*>(p.preact_out + static_cast<int64_t>(mw) * p.ld_preact + n0), p.ld_preact * 2,
               rows, false);
  }
  if (has<EPI, EF_OUT_BF16>(p.out_bf16 != nullptr)) {
    tile_put(sa, lane, wout);
    tile_store(sa, lane, reinterpret_cast<uint8_t*>(p.out_bf16 + static_cast<int64_t>(mw) * p.ld_out_bf16 + n0), p.ld_out_bf16 * 2,
               rows, false);
  }
}

// L2 prefetch of the per-element epilogue INPUTS (fp32 residual, bf16 aux, fp32 accumulate target) of this warp's part
// of its NEXT tile: lane l touches the 128-byte lines of row mw + l.  Issued one tile ahead, so the loads inside the
// chunk epilogue hit L2 (~250 cycles) instead of paying a loaded-DRAM round trip on every chunk's critical path (ncu:
// the top stall of the "+ residual" epilogues was the st.shared that waits for the residual load).
template <int EPI, int COLS>
__device__ __forceinline__ void epilogue_prefetch(const EpiParams& p, int lane, int mw, int n_first) {
  const int m = mw + lane;
  if (m >= p.M || n_first >= p.N) return;
  const int cols = min(COLS, p.N - n_first);
  if (has<EPI, EF_RESIDUAL>(p.residual != nullptr)) {
    const char* a = reinterpret_cast<const char*>(p.residual + static_cast<int64_t>(m) * p.ld_res + n_first);
    for (int b = 0; b < cols * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + b));
  }
  if (has<EPI, EF_MUL_AUX>(p.act == X2K_ACT_MUL_AUX) || has<EPI, EF_GELU_BWD>(p.act == X2K_ACT_GELU_BWD)) {
    const char* a = reinterpret_cast<const char*>(p.aux + static_cast<int64_t>(m) * p.ld_aux + n_first);
    for (int b = 0; b < cols * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + b));
  }
  if (has<EPI, EF_OUT_F32>(p.out_f32 != nullptr) && p.accumulate && p.split_k <= 1) {
    const char* a = reinterpret_cast<const char*>(p.out_f32 + static_cast<int64_t>(m) * p.ld_out_f32 + n_first);
    for (int b = 0; b < cols * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + b));
  }
}

// The 32-column chunks of one warp: full chunks through the coalesced path, a ragged last chunk (N % 32 != 0) through
// the thread-per-row path.  Warp-uniform control flow around every tcgen05.ld.
template <int EPI, int NCH>
__device__ __forceinline__ void epilogue_warp_columns(const EpiParams& p, uint32_t sa, int lane, int mw, int n_first,
                                                      uint32_t taddr) {
  const int m = mw + lane;
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    const int n = n_first + ci * 32;
    if (mw < p.M && n < p.N) {
      if (n + 32 <= p.N) {
        epilogue_chunk32<EPI>(p, sa, lane, mw, min(32, p.M - mw), n, taddr + ci * 32);
      } else {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (n + 16 * hf < p.N) {
            uint32_t acc[16];
            tmem_ld_32x16(taddr + ci * 32 + 16 * hf, acc);
            tmem_wait_ld();
            if (m < p.M) epilogue_tail16(p, m, n + 16 * hf, acc);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA-staged epilogue of the fp32 residual-stream GEMMs (bias [+ pre-activation] [+ dropout] [+ LayerScale / DropPath]
// + residual -> fp32): out = f(acc) + residual.  The thread-per-row <-> coalesced transposition still goes through the
// per-warp 32 x 64-byte tile, but the tile's GLOBAL side is moved by the TMA engine instead of LDG/STS and LDS/STG:
//   * the residual tile of the NEXT 16-column half is requested (cp.async.bulk.tensor, 64B swizzle == epi_off layout)
//     as soon as the current one has been read, so its latency hides behind the TMEM load / bias / dropout math of that
//     half — and the first half of the next output tile is requested before this warp goes to wait for the MMAs;
//   * the finished fp32 half is written into a second tile and leaves by one cp.async.bulk.tensor store
//     (UTMASTG): no per-thread global stores, out-of-range rows are clipped by the tensor map.
// One 2 KB input tile + one 2 KB output tile + one mbarrier per epilogue warp (5 operand stages instead of 6).
// ---------------------------------------------------------------------------------------------
template <int EPI>
constexpr bool kTmaEpi = EPI != EF_GENERIC && (EPI & EF_RESIDUAL) != 0 && (EPI & EF_OUT_F32) != 0;

struct TmaEpiState {
  uint32_t sa_r, sa_o;  // shared-space addresses of the residual (in) and output tiles of this warp
  uint64_t* bar;        // completion barrier of the residual loads
  uint32_t phase;
  bool pending;         // a residual load is in flight
};

template <int EPI>
__device__ __forceinline__ void epilogue_warp_tma(const EpiParams& p, const CUtensorMap* tm_res, const CUtensorMap* tm_out,
                                                  TmaEpiState& st, int lane, int mw, int n_first, uint32_t taddr, int next_mw,
                                                  int next_n) {
  constexpr int NCH = 2;  // 64 columns per warp
  const int m = mw + lane;
  const bool rows_live = mw < p.M;
  const int rows = min(32, p.M - mw);
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    const int n0 = n_first + ci * 32;
    const bool live = rows_live && n0 < p.N;  // N % 32 == 0 on this path: a chunk is whole or absent
    uint32_t wpre[16];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int nh = n0 + 16 * hf;
      float v[16];
      if (live) {
        uint32_t acc[16];
        tmem_ld_32x16(taddr + ci * 32 + 16 * hf, acc);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        if constexpr ((EPI & EF_BIAS) != 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nh + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if constexpr ((EPI & EF_PREACT) != 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) wpre[8 * hf + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        }
        if constexpr ((EPI & EF_DROPOUT) != 0) {
          if (p.dropout_p > 0.0f) {
            const DropCfg dc = make_drop(p.dropout_p);
            const uint64_t base = static_cast<uint64_t>(m) * static_cast<uint64_t>(p.N) + static_cast<uint64_t>(nh);
            const uint64_t doff = p.dropout_offset + (p.dropout_offset_dev ? __ldg(p.dropout_offset_dev) : 0ull);
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              float k[8];
              drop8(p.dropout_seed, doff, (base + j) >> 3, dc, k);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[j + i] *= k[i];
            }
          }
        }
        if constexpr ((EPI & EF_SCALE) != 0) {
          if (p.gamma) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + nh + j));
              v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
            }
          }
          if (p.row_scale) {
            const float sc = __ldg(p.row_scale + min(m, p.M - 1) / p.rows_per_scale);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= sc;
          }
        }
      }
      // ---- residual half: landed by TMA while the math above ran ----
      if (st.pending) {  // warp-uniform
        mbar_wait_warp(st.bar, st.phase);
        st.phase ^= 1;
        uint32_t w[16];
        tile_get(st.sa_r, lane, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(w[j]);  // consumes the loads before the tile is re-filled
        __syncwarp();
      }
      // ---- request the next half this warp will need: same chunk, next chunk, or the next output tile ----
      {
        int pn, pm;
        bool pv;
        if (hf == 0) { pn = nh + 16; pm = mw; pv = live; }
        else if (ci + 1 < NCH) { pn = n0 + 32; pm = mw; pv = rows_live && pn < p.N; }
        else { pn = next_n; pm = next_mw; pv = next_mw >= 0 && next_mw < p.M && next_n < p.N; }
        if (pv && lane == 0) {
          mbar_arrive_expect_tx(st.bar, EPI_TILE_BYTES);
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(st.sa_r), "l"(reinterpret_cast<uint64_t>(tm_res)), "r"(smem_u32(st.bar)), "r"(pn), "r"(pm) : "memory");
        }
        st.pending = pv;
      }
      // ---- fp32 half -> output tile -> one TMA store ----
      if (live) {
        if (lane == 0) tma_store_wait_read();  // the previous store has finished reading the output tile
        __syncwarp();
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(v[j]);
        tile_put(st.sa_o, lane, w);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tm_out, st.sa_o, nh, mw);
          tma_store_commit();
        }
      }
    }
    if constexpr ((EPI & EF_PREACT) != 0) {
      if (live) {  // bf16 pre-activation of the whole 32-column chunk through the output tile (synchronous path)
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        tile_put(st.sa_o, lane, wpre);
        tile_store(st.sa_o, lane, reinterpret_cast<uint8_t*>(p.preact_out + static_cast<int64_t>(mw) * p.ld_preact + n0),
                   p.ld_preact * 2, rows, false);
      }
    }
  }
}

template <int BLOCK_N, int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const EpiParams p) {
  using C = Cfg<BLOCK_N>;
  constexpr int STAGES = C::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t epi_stage = smem_u32(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + C::EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int k_blocks_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int split_k = p.split_k > 1 ? p.split_k : 1;
  const int kb_per = (k_blocks_total + split_k - 1) / split_k;
  const int num_tiles = m_tiles * n_tiles * split_k;  // work units: unit = tile * split_k + k-slice

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], NUM_EPI_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int t = tile / split_k, ks = tile % split_k;
        const int m0 = (p.raster_m ? t % m_tiles : t / n_tiles) * BLOCK_M;
        const int n0 = (p.raster_m ? t / m_tiles : t % n_tiles) * BLOCK_N;
        const int kb_end = min(k_blocks_total, (ks + 1) * kb_per);
        for (int kb = ks * kb_per; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if (A_MN == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + c * 64, k0);
          }
          if (B_MN == 0) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_N / 64; ++c)
              tma_load_2d(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + c * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; the other 31 lanes wait at the final barrier) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * BLOCK_N;
        const int ks = tile % split_k;
        const int kb_begin = ks * kb_per, kb_end = min(k_blocks_total, (ks + 1) * kb_per);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = A_MN == 0 ? make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024)
                                          : make_smem_desc(sa + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
            const uint64_t db = B_MN == 0 ? make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024)
                                          : make_smem_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == kb_end - 1) umma_commit(&tmem_full_bar[acc_stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;            // 0..15
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int part = ew >> 2;           // which quarter of the BLOCK_N columns
    constexpr int COLS_PER_WARP = BLOCK_N / (NUM_EPI_WARPS / 4);
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int t = tile / split_k;
      const int m0 = (p.raster_m ? t % m_tiles : t / n_tiles) * BLOCK_M;
      const int n0 = (p.raster_m ? t / m_tiles : t % n_tiles) * BLOCK_N;
      mbar_wait_warp(&tmem_full_bar[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc_stage * BLOCK_N + part * COLS_PER_WARP + (static_cast<uint32_t>(quad * 32) << 16);
      if (tile + static_cast<int>(gridDim.x) < num_tiles) {  // inputs of the next tile -> L2
        const int tn = (tile + static_cast<int>(gridDim.x)) / split_k;
        const int m0n = (p.raster_m ? tn % m_tiles : tn / n_tiles) * BLOCK_M;
        const int n0n = (p.raster_m ? tn / m_tiles : tn % n_tiles) * BLOCK_N;
        epilogue_prefetch<EPI, COLS_PER_WARP>(p, lane, m0n + quad * 32, n0n + part * COLS_PER_WARP);
      }
      epilogue_warp_columns<EPI, COLS_PER_WARP / 32>(p, epi_stage + ew * EPI_TILE_BYTES, lane, m0 + quad * 32,
                                                     n0 + part * COLS_PER_WARP, taddr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc_stage]);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  A cluster of two CTAs on one TPC computes a 256 x 256 tile: CTA r holds rows
// [128r, 128r+128) of A and of the accumulator, and HALF of the B tile (128 of the 256 output columns); one
// tcgen05.mma.cta_group::2 issued by CTA 0 drives both tensor cores, each reading the B halves of both CTAs.  Per SM
// and per k-block that is 32 KB of TMA writes and 8 KB of operand reads per MMA instead of 48 KB / 12 KB — the 1-CTA
// kernel is bound by shared-memory bandwidth (TMA fill + operand fetch > 128 B/clk), this one is not.
// Barriers: full[] lives in CTA 0 and collects the TMA bytes of both CTAs; empty[] / tmem_full[] are per CTA and are
// signalled by multicast tcgen05.commit; tmem_empty[] lives in CTA 0 and is arrived on by the epilogue warps of both.
// ---------------------------------------------------------------------------------------------
template <int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                         const EpiParams p) {
  constexpr int BLOCK_N = 256;
  constexpr int HALF_N = 128;
  constexpr int STAGE_BYTES = A_TILE_BYTES + HALF_N * BLOCK_K * 2;  // 32 KB
  // 6 x 32 KB of operand stages + 32 KB of epilogue transpose tiles — or, for the fp32 residual-stream epilogues, 5 stages
  // + 64 KB: every epilogue warp gets a TMA-filled residual tile next to its output tile (epilogue_warp_tma)
  constexpr bool TMA_EPI = kTmaEpi<EPI>;
  constexpr int STAGES = TMA_EPI ? 5 : 6;
  constexpr int EPI_BYTES_PER_WARP = TMA_EPI ? 2 * EPI_TILE_BYTES : EPI_TILE_BYTES;
  constexpr int TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t epi_stage = smem_u32(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + NUM_EPI_WARPS * EPI_BYTES_PER_WARP);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* epi_bar = tmem_empty_bar + 2;  // [NUM_EPI_WARPS], TMA_EPI only
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(epi_bar + (TMA_EPI ? NUM_EPI_WARPS : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());  // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int k_blocks_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int split_k = p.split_k > 1 ? p.split_k : 1;
  const int kb_per = (k_blocks_total + split_k - 1) / split_k;
  const int num_tiles = m_tiles * n_tiles * split_k;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 2);   // CTA 0: own arrive.expect_tx + the peer's arrive
        mbar_init(&empty_bar[s], 1);  // multicast commit of the leader's MMA thread
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full_bar[s], 1);
        mbar_init(&tmem_empty_bar[s], 2 * NUM_EPI_WARPS);  // CTA 0: epilogue warps of both CTAs
      }
      if (TMA_EPI) {
        for (int s = 0; s < NUM_EPI_WARPS; ++s) mbar_init(&epi_bar[s], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_ptr, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int t = tile / split_k, ks = tile % split_k;
        const int m0 = (t / n_tiles) * (2 * BLOCK_M) + rank * BLOCK_M;
        const int n0 = (t % n_tiles) * BLOCK_N + rank * HALF_N;
        const int kb_end = min(k_blocks_total, (ks + 1) * kb_per);
        for (int kb = ks * kb_per; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          const int k0 = kb * BLOCK_K;
          if (A_MN == 0) {
            tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d_pair(sa + c * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + c * 64, k0);
          }
          if (B_MN == 0) {
            tma_load_2d_pair(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < HALF_N / 64; ++c)
              tma_load_2d_pair(sb + c * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + c * 64, k0);
          }
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          else mbar_arrive_cta0(&full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * BLOCK_N;
        const int ks = tile % split_k;
        const int kb_begin = ks * kb_per, kb_end = min(k_blocks_total, (ks + 1) * kb_per);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = A_MN == 0 ? make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024)
                                          : make_smem_desc(sa + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
            const uint64_t db = B_MN == 0 ? make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024)
                                          : make_smem_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024);
            umma_bf16_pair(d_tmem, da, db, idesc, (kb != kb_begin || k != 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], 3);  // frees this smem slot in both CTAs
          if (kb == kb_end - 1) umma_commit_pair(&tmem_full_bar[acc_stage], 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, each on its own 128 rows) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int part = ew >> 2;
    constexpr int COLS_PER_WARP = BLOCK_N / (NUM_EPI_WARPS / 4);
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    const bool use_tma = TMA_EPI && p.tma_epi;
    TmaEpiState st;
    st.sa_r = epi_stage + ew * EPI_BYTES_PER_WARP;
    st.sa_o = st.sa_r + EPI_TILE_BYTES;
    st.bar = &epi_bar[TMA_EPI ? ew : 0];
    st.phase = 0;
    st.pending = false;
    if constexpr (TMA_EPI) {
      if (use_tma && pair < num_tiles) {  // residual of this warp's first half: lands while the first tile's MMAs run
        const int t0 = pair / split_k;
        const int pm = (t0 / n_tiles) * (2 * BLOCK_M) + rank * BLOCK_M + quad * 32, pn = (t0 % n_tiles) * BLOCK_N + part * COLS_PER_WARP;
        const bool pv = pm < p.M && pn < p.N;
        if (pv && lane == 0) {
          mbar_arrive_expect_tx(st.bar, EPI_TILE_BYTES);
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(st.sa_r), "l"(reinterpret_cast<uint64_t>(&tmap_res)), "r"(smem_u32(st.bar)), "r"(pn), "r"(pm) : "memory");
        }
        st.pending = pv;
      }
    }
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int t = tile / split_k;
      const int m0 = (t / n_tiles) * (2 * BLOCK_M) + rank * BLOCK_M;
      const int n0 = (t % n_tiles) * BLOCK_N;
      mbar_wait_warp(&tmem_full_bar[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc_stage * BLOCK_N + part * COLS_PER_WARP + (static_cast<uint32_t>(quad * 32) << 16);
      const bool has_next = tile + num_pairs < num_tiles;
      const int tn = has_next ? (tile + num_pairs) / split_k : 0;
      const int m0n = (tn / n_tiles) * (2 * BLOCK_M) + rank * BLOCK_M, n0n = (tn % n_tiles) * BLOCK_N;
      bool done = false;
      if constexpr (TMA_EPI) {
        if (use_tma) {
          if (has_next)  // next tile's residual rows -> L2, so the per-half TMA loads only pay an L2 hit
            epilogue_prefetch<EPI, COLS_PER_WARP>(p, lane, m0n + quad * 32, n0n + part * COLS_PER_WARP);
          epilogue_warp_tma<EPI>(p, &tmap_res, &tmap_out, st, lane, m0 + quad * 32, n0 + part * COLS_PER_WARP, taddr,
                                 has_next ? m0n + quad * 32 : -1, n0n + part * COLS_PER_WARP);
          done = true;
        }
      }
      if (!done) {
        if (has_next)  // inputs of the next tile -> L2
          epilogue_prefetch<EPI, COLS_PER_WARP>(p, lane, m0n + quad * 32, n0n + part * COLS_PER_WARP);
        epilogue_warp_columns<EPI, COLS_PER_WARP / 32>(p, epi_stage + ew * EPI_BYTES_PER_WARP, lane, m0 + quad * 32,
                                                       n0 + part * COLS_PER_WARP, taddr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta0(&tmem_empty_bar[acc_stage]);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
    if constexpr (TMA_EPI) {
      if (use_tma && lane == 0) tma_store_wait_read();  // the last store has read its tile before the CTA may exit
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its partner can still signal or read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// 6 stages + 16 x 2 KB tiles, or 5 stages + 16 x 4 KB tiles (TMA-staged epilogue): the same 224 KB; barriers: 2*STAGES + 4 (+ 16)
constexpr int PAIR_SMEM_BYTES = 6 * (A_TILE_BYTES + 128 * BLOCK_K * 2) + NUM_EPI_WARPS * EPI_TILE_BYTES + 1024 + 256;
static_assert(5 * (A_TILE_BYTES + 128 * BLOCK_K * 2) + NUM_EPI_WARPS * 2 * EPI_TILE_BYTES == 6 * (A_TILE_BYTES + 128 * BLOCK_K * 2) + NUM_EPI_WARPS * EPI_TILE_BYTES, "smem budget");
static_assert((2 * 5 + 4 + NUM_EPI_WARPS) * 8 + 8 <= 256 && (2 * 6 + 4) * 8 + 8 <= 256, "barrier area");

template <int A_MN, int B_MN, int EPI>
int launch_pair(const X2kGemmArgs& a, const EpiParams& ep, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc;
  if (A_MN == 0) rc = make_tmap_bf16_2d(&ta, a.A, a.M, a.K, a.lda, BLOCK_M, BLOCK_K);
  else rc = make_tmap_bf16_2d(&ta, a.A, a.K, a.M, a.lda, BLOCK_K, 64);
  if (rc) return rc;
  if (B_MN == 0) rc = make_tmap_bf16_2d(&tb, a.B, a.N, a.K, a.ldb, 128, BLOCK_K);
  else rc = make_tmap_bf16_2d(&tb, a.B, a.K, a.N, a.ldb, BLOCK_K, 64);
  if (rc) return rc;
  CUtensorMap tres = ta, tout = ta;  // placeholders unless the TMA-staged epilogue runs
  if (kTmaEpi<EPI> && ep.tma_epi) {
    if ((rc = make_tmap_f32_epi(&tres, a.residual, a.M, a.N, a.ld_res))) return rc;
    if ((rc = make_tmap_f32_epi(&tout, a.out_f32, a.M, a.N, a.ld_out_f32))) return rc;
  }
  auto kern = gemm_tcgen05_pair_kernel<A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (a.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (a.N + 255) / 256;
  const int units = m_tiles * n_tiles * (ep.split_k > 1 ? ep.split_k : 1);
  int pairs = sm_count() / 2;
  if (a.max_ctas > 0 && a.max_ctas / 2 < pairs) pairs = a.max_ctas / 2 > 0 ? a.max_ctas / 2 : 1;
  if (units < pairs) pairs = units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = PAIR_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  X2K_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tres, tout, ep));
  count_launch();
  return X2K_OK;
}

template <int BLOCK_N, int A_MN, int B_MN, int EPI>
int launch(const X2kGemmArgs& a, const EpiParams& ep, cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  CUtensorMap ta, tb;
  int rc;
  if (A_MN == 0)
    rc = make_tmap_bf16_2d(&ta, a.A, a.M, a.K, a.lda, BLOCK_M, BLOCK_K);
  else
    rc = make_tmap_bf16_2d(&ta, a.A, a.K, a.M, a.lda, BLOCK_K, 64);
  if (rc) return rc;
  if (B_MN == 0)
    rc = make_tmap_bf16_2d(&tb, a.B, a.N, a.K, a.ldb, BLOCK_N, BLOCK_K);
  else
    rc = make_tmap_bf16_2d(&tb, a.B, a.K, a.N, a.ldb, BLOCK_K, 64);
  if (rc) return rc;

  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, EPI>;
  static bool attr_set = false;  // idempotent; racing writers set the same value
  if (!attr_set) {
    X2K_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (a.N + BLOCK_N - 1) / BLOCK_N;
  int ctas = sm_count();
  if (a.max_ctas > 0 && a.max_ctas < ctas) ctas = a.max_ctas;
  const int units = m_tiles * n_tiles * (ep.split_k > 1 ? ep.split_k : 1);
  if (units < ctas) ctas = units;
  kern<<<ctas, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, ep);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

template <int BLOCK_N, int EPI>
int dispatch_major(const X2kGemmArgs& a, const EpiParams& ep, cudaStream_t stream) {
  if (BLOCK_N == 256 && ep.use_pair) {
    if (!a.a_mn_major && !a.b_mn_major) return launch_pair<0, 0, EPI>(a, ep, stream);
    if (!a.a_mn_major && a.b_mn_major) return launch_pair<0, 1, EPI>(a, ep, stream);
    if (a.a_mn_major && !a.b_mn_major) return launch_pair<1, 0, EPI>(a, ep, stream);
    return launch_pair<1, 1, EPI>(a, ep, stream);
  }
  if (!a.a_mn_major && !a.b_mn_major) return launch<BLOCK_N, 0, 0, EPI>(a, ep, stream);
  if (!a.a_mn_major && a.b_mn_major) return launch<BLOCK_N, 0, 1, EPI>(a, ep, stream);
  if (a.a_mn_major && !a.b_mn_major) return launch<BLOCK_N, 1, 0, EPI>(a, ep, stream);
  return launch<BLOCK_N, 1, 1, EPI>(a, ep, stream);
}

// feature mask of a call (what the runtime pointers ask for)
int epi_mask(const X2kGemmArgs& a) {
  int m = 0;
  if (a.bias) m |= EF_BIAS;
  if (a.preact_out) m |= EF_PREACT;
  if (a.act == X2K_ACT_GELU) m |= EF_GELU;
  if (a.act == X2K_ACT_GELU_BWD) m |= EF_GELU_BWD;
  if (a.act == X2K_ACT_GELU_SAVE_GRAD) m |= EF_GELU_SAVE;
  if (a.act == X2K_ACT_MUL_AUX) m |= EF_MUL_AUX;
  if (a.dropout_p > 0.f) m |= EF_DROPOUT;
  if (a.gamma || a.row_scale) m |= EF_SCALE;
  if (a.residual) m |= EF_RESIDUAL;
  if (a.out_f32) m |= EF_OUT_F32;
  if (a.out_bf16) m |= EF_OUT_BF16;
  return m;
}

// epilogue variants: index -> compile-time feature mask.  A specialised variant may contain MORE features than requested
// only where the extra feature is itself guarded by a runtime test (dropout_p, gamma/row_scale pointers), never fewer.
#define X2K_EPI_LIST(X)                                                                   \
  X(0, EF_OUT_BF16)                                               /* plain dgrad */                        \
  X(1, EF_BIAS | EF_OUT_BF16)                                     /* qkv / q / kv projections */           \
  X(2, EF_BIAS | EF_PREACT | EF_GELU_SAVE | EF_OUT_BF16)          /* fc1 / intermediate: GELU, GELU' saved */ \
  X(3, EF_BIAS | EF_PREACT | EF_SCALE | EF_RESIDUAL | EF_OUT_F32) /* BEiT proj / fc2 (LayerScale, DropPath) */ \
  X(4, EF_BIAS | EF_PREACT | EF_RESIDUAL | EF_OUT_F32)            /* same without LayerScale / DropPath */ \
  X(5, EF_BIAS | EF_DROPOUT | EF_RESIDUAL | EF_OUT_F32)           /* BERT dense + dropout + residual */    \
  X(6, EF_BIAS | EF_RESIDUAL | EF_OUT_F32)                        /* BERT dense + residual (eval) */       \
  X(7, EF_MUL_AUX | EF_OUT_BF16)                                  /* dgrad through GELU (saved GELU') */    \
  X(8, EF_OUT_F32)                                                /* wgrad (accumulate / split-K atomics) */ \
  X(9, EF_RESIDUAL | EF_OUT_F32)                                  /* dgrad + residual-stream gradient */   \
  X(10, EF_GENERIC)                                               /* anything else: runtime tests */

}  // namespace
#ifndef X2K_GEMM_PART
#error "gemm.cu is compiled in parts: pass -DX2K_GEMM_PART=0..3 (x2vlm_b200/build.py does)"
#endif
#define X2K_PART_FN_(n) gemm_epi_part##n
#define X2K_PART_FN(n) X2K_PART_FN_(n)
int gemm_epi_part0(int, int, const X2kGemmArgs&, const EpiParams&, cudaStream_t);
int gemm_epi_part1(int, int, const X2kGemmArgs&, const EpiParams&, cudaStream_t);
int gemm_epi_part2(int, int, const X2kGemmArgs&, const EpiParams&, cudaStream_t);
int gemm_epi_part3(int, int, const X2kGemmArgs&, const EpiParams&, cudaStream_t);

// this translation unit's share of the variants
int X2K_PART_FN(X2K_GEMM_PART)(int idx, int tile_n, const X2kGemmArgs& a, const EpiParams& ep, cudaStream_t stream) {
  switch (idx) {
#define X2K_EPI_CASE(I, MASK)                                                                                  \
  case I:                                                                                                      \
    if constexpr ((I) % 4 == X2K_GEMM_PART)                                                                    \
      return tile_n == 256 ? dispatch_major<256, (MASK)>(a, ep, stream) : dispatch_major<128, (MASK)>(a, ep, stream); \
    break;
    X2K_EPI_LIST(X2K_EPI_CASE)
#undef X2K_EPI_CASE
  }
  set_error("x2k_gemm: epilogue variant %d is not in part %d", idx, X2K_GEMM_PART);
  return X2K_ERR_UNSUPPORTED;
}
namespace {

#if X2K_GEMM_PART == 0
int dispatch_epi(int tile_n, const X2kGemmArgs& a, const EpiParams& ep, cudaStream_t stream) {
  const int m = epi_mask(a);
  int idx = 10;
#define X2K_EPI_CASE(I, MASK) \
  if ((MASK) != EF_GENERIC && m == (MASK)) idx = I;
  X2K_EPI_LIST(X2K_EPI_CASE)
#undef X2K_EPI_CASE
  if (a.ce_mode) idx = 10;  // the fused cross entropy lives in the generic variant
  switch (idx % 4) {
    case 0: return gemm_epi_part0(idx, tile_n, a, ep, stream);
    case 1: return gemm_epi_part1(idx, tile_n, a, ep, stream);
    case 2: return gemm_epi_part2(idx, tile_n, a, ep, stream);
    default: return gemm_epi_part3(idx, tile_n, a, ep, stream);
  }
}
#endif

}  // namespace
}  // namespace x2k

#if X2K_GEMM_PART == 0
extern "C" int x2k_gemm(const X2kGemmArgs* args, void* stream_) {
  using namespace x2k;
  X2K_REQUIRE(args != nullptr, "x2k_gemm: args is NULL");
  const X2kGemmArgs& a = *args;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(a.A && a.B, "x2k_gemm: A/B is NULL");
  X2K_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "x2k_gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  X2K_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "x2k_gemm: lda/ldb must be multiples of 8 (got %lld, %lld)",
              (long long)a.lda, (long long)a.ldb);
  X2K_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
              "x2k_gemm: A/B must be 16-byte aligned");
  X2K_REQUIRE(a.out_bf16 || a.out_f32 || a.preact_out || a.ce_mode == 1, "x2k_gemm: no output");
  X2K_REQUIRE(a.ce_mode >= 0 && a.ce_mode <= 2, "x2k_gemm: ce_mode must be 0, 1 or 2");
  X2K_REQUIRE(a.ce_mode != 1 || (a.ce_partials && a.ce_target_logit && a.ce_labels), "x2k_gemm: ce_mode 1 needs ce_partials, ce_target_logit, ce_labels");
  X2K_REQUIRE(a.ce_mode != 2 || (a.ce_lse && a.ce_row_grad && a.ce_labels && a.out_bf16), "x2k_gemm: ce_mode 2 needs ce_lse, ce_row_grad, ce_labels, out_bf16");
  X2K_REQUIRE(!a.ce_mode || a.act == X2K_ACT_NONE, "x2k_gemm: the fused cross entropy takes no activation");
  X2K_REQUIRE(!a.accumulate || a.out_f32, "x2k_gemm: accumulate needs out_f32");
  X2K_REQUIRE(a.act >= X2K_ACT_NONE && a.act <= X2K_ACT_MUL_AUX, "x2k_gemm: unknown act %d", a.act);
  X2K_REQUIRE((a.act != X2K_ACT_GELU_BWD && a.act != X2K_ACT_MUL_AUX) || a.aux, "x2k_gemm: GELU_BWD / MUL_AUX need aux");
  X2K_REQUIRE(a.act != X2K_ACT_GELU_SAVE_GRAD || a.preact_out, "x2k_gemm: GELU_SAVE_GRAD needs preact_out");
  X2K_REQUIRE(!(a.dropout_p > 0.f) || (a.N % 8 == 0 && a.dropout_p < 1.f), "x2k_gemm: dropout needs N%%8==0, p<1");
  X2K_REQUIRE(!a.row_scale || a.rows_per_scale > 0, "x2k_gemm: rows_per_scale must be > 0");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  X2K_REQUIRE(al16(a.bias) && al16(a.gamma) && al16(a.residual) && al16(a.out_f32) && al16(a.out_bf16) &&
                  al16(a.preact_out) && al16(a.aux),
              "x2k_gemm: epilogue pointers must be 16-byte aligned");
  X2K_REQUIRE((!a.out_bf16 || a.ld_out_bf16 % 8 == 0) && (!a.out_f32 || a.ld_out_f32 % 4 == 0) &&
                  (!a.preact_out || a.ld_preact % 8 == 0) && (!a.aux || a.ld_aux % 8 == 0) &&
                  (!a.residual || a.ld_res % 4 == 0),
              "x2k_gemm: epilogue leading dimensions must keep 16-byte row alignment");

  EpiParams ep;
  ep.M = a.M; ep.N = a.N; ep.K = a.K;
  ep.bias = a.bias; ep.act = a.act;
  ep.aux = static_cast<const __nv_bfloat16*>(a.aux); ep.ld_aux = a.ld_aux;
  ep.preact_out = static_cast<__nv_bfloat16*>(a.preact_out); ep.ld_preact = a.ld_preact;
  ep.dropout_p = a.dropout_p; ep.dropout_seed = a.dropout_seed; ep.dropout_offset = a.dropout_offset;
  ep.dropout_offset_dev = a.dropout_offset_dev;
  ep.gamma = a.gamma; ep.row_scale = a.row_scale; ep.rows_per_scale = a.rows_per_scale;
  ep.residual = a.residual; ep.ld_res = a.ld_res; ep.accumulate = a.accumulate;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(a.out_bf16); ep.ld_out_bf16 = a.ld_out_bf16;
  ep.out_f32 = a.out_f32; ep.ld_out_f32 = a.ld_out_f32;
  ep.dbg_lbo = ep.dbg_sbo = ep.dbg_kadv = 0;
  ep.ce_mode = a.ce_mode; ep.ce_labels = a.ce_labels; ep.ce_partials = a.ce_partials; ep.ce_target_logit = a.ce_target_logit;
  ep.ce_lse = a.ce_lse; ep.ce_row_grad = a.ce_row_grad;
  ep.raster_m = 0;
  if (const char* r = getenv("X2K_GEMM_RASTER")) ep.raster_m = atoi(r);

  // ---- tiling: CTA-pair kernel (256 x 256 cluster tiles) whenever the output has >= 256 rows and columns, else the
  //      1-CTA kernel with the tile width that minimises (waves x tile cost) ----
  int tile_n = a.tile_n;
  ep.use_pair = 0;
  {
    int mode = 2;  // 0 = never, 1 = whenever the tile is 256 wide, 2 = auto
    if (const char* e = getenv("X2K_GEMM_PAIR")) mode = atoi(e);
    const bool can = (a.tile_n == 0 || a.tile_n == 256) && a.N >= 256;
    if (can && (mode == 1 || (mode == 2 && a.tile_n == 0 && a.M >= 256))) { ep.use_pair = 1; tile_n = 256; }
  }
  // ---- split-K: a wgrad-shaped GEMM (small output, very long K) has fewer work units than SMs; slice K across CTAs
  //      and accumulate the fp32 output with vector atomics.  Only for the pure (accumulating) fp32 epilogue. ----
  ep.split_k = 1;
  const bool plain_f32 = a.out_f32 && !a.out_bf16 && !a.preact_out && !a.bias && a.act == X2K_ACT_NONE &&
                         !(a.dropout_p > 0.f) && !a.gamma && !a.row_scale && !a.residual;
  if (plain_f32 && a.split_k != 1) {
    const int sms = sm_count();
    // work units and the number of them that run concurrently with the chosen tiling
    const long units = ep.use_pair ? static_cast<long>((a.M + 255) / 256) * ((a.N + 255) / 256)
                                   : static_cast<long>((a.M + BLOCK_M - 1) / BLOCK_M) * ((a.N + 127) / 128);
    const long slots = ep.use_pair ? sms / 2 : sms;
    const int kbt = (a.K + BLOCK_K - 1) / BLOCK_K;
    int s = a.split_k > 1 ? a.split_k : 1;
    if (a.split_k == 0 && units < slots && kbt >= 32) {
      // Work units are dealt round-robin to the persistent CTAs (pairs), so the launch takes ceil(units * s / slots)
      // rounds of one slice each: pick the largest s whose units still fit WHOLE rounds (2 by default — the second
      // round's MMAs hide the first round's atomic epilogue).  Rounding s up instead (17 slices of a 9-tile output on 74
      // pairs = 153 units) spills a few units into a third round and costs ~30% of the launch.
      static const int waves = getenv("X2K_GEMM_SPLITK_WAVES") ? atoi(getenv("X2K_GEMM_SPLITK_WAVES")) : 2;
      s = static_cast<int>((static_cast<long>(waves > 0 ? waves : 2) * slots) / units);
      if (s > kbt / 8) s = kbt / 8;  // keep >= 8 k-blocks per slice
      if (s < 1) s = 1;
    }
    if (s > 1) {  // no slice may be empty (an MMA warp without k-blocks would never signal its epilogue)
      const int per = (kbt + s - 1) / s;
      s = (kbt + per - 1) / per;
    }
    if (s > 1) {
      ep.split_k = s;
      if (!a.accumulate)  // atomics need a defined starting value
        X2K_CHECK_CUDA(cudaMemset2DAsync(a.out_f32, a.ld_out_f32 * sizeof(float), 0, static_cast<size_t>(a.N) * sizeof(float),
                                         a.M, stream));
    }
  }
  // TMA-staged fp32 residual epilogue (pair kernel): whole 32-column chunks only, plain (non-accumulating) fp32 output
  ep.tma_epi = 0;
  {
    static const bool off = getenv("X2K_GEMM_NO_TMA_EPI") != nullptr;  // developer switch for A/B timing
    if (!off && ep.use_pair && a.residual && a.out_f32 && !a.out_bf16 && !a.accumulate && ep.split_k <= 1 && a.N % 32 == 0 &&
        a.ce_mode == 0)
      ep.tma_epi = 1;
  }
  if (!ep.use_pair) {
    if (ep.split_k > 1 && tile_n == 0) tile_n = 128;
    if (tile_n == 0) {
      const int sms = sm_count();
      const long m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
      const long t256 = m_tiles * ((a.N + 255) / 256), t128 = m_tiles * ((a.N + 127) / 128);
      const long cost256 = ((t256 + sms - 1) / sms) * 2, cost128 = ((t128 + sms - 1) / sms) * 1;
      tile_n = (a.N <= 128 || cost128 < cost256) ? 128 : 256;
    }
  }
  X2K_REQUIRE(tile_n == 128 || tile_n == 256, "x2k_gemm: tile_n must be 0, 128 or 256");
  return dispatch_epi(tile_n, a, ep, stream);
}
#endif  // X2K_GEMM_PART == 0
