// rowops.cu — HBM-bound kernels of the hot path: LayerNorm fwd/bwd, epilogue backward
// (scale/cast/column-sum), bias-gradient column sums, casts, flat AdamW.  All are streaming
// kernels: 16-byte vectorised, coalesced along the contiguous dimension, grids sized as multiples
// of the SM count; reductions are warp-shuffle -> smem -> few atomics.
#include "common.cuh"

namespace x2k {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, row cached in registers (D = 128 * VEC, VEC <= 8).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, int M, int D, float eps,
                                                            __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * D);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    v[j] = xr[lane + 32 * j];
    s += v[j].x + v[j].y + v[j].z + v[j].w;
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
    q += a * a + c * c + d * d + e * e;
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int c4 = lane + 32 * j;
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c4);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + c4);
    float4 o;
    o.x = (v[j].x - mean) * rstd * ww.x + bb.x;
    o.y = (v[j].y - mean) * rstd * ww.y + bb.y;
    o.z = (v[j].z - mean) * rstd * ww.z + bb.z;
    o.w = (v[j].w - mean) * rstd * ww.w + bb.w;
    if (y_f32) reinterpret_cast<float4*>(y_f32 + static_cast<int64_t>(row) * D)[c4] = o;
    if (y_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y_bf16 + static_cast<int64_t>(row) * D)[c4] = pk;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward: warps stride over rows, keep per-column dw/db partials in registers,
// reduce across the block's warps in smem, then one atomicAdd per column per block.
// ---------------------------------------------------------------------------------------------
// FUSE: the kernel also produces what x2k_scale_cast_colsum would compute from dx in a second pass over it (the BERT
// "dense + dropout + residual -> LayerNorm" backward): g = bf16(dx * keepscale) and dbias += colsum(dx * keepscale).
struct LnFuse {
  float dropout_p;
  uint64_t seed, offset;
  const uint64_t* offset_dev;
  __nv_bfloat16* g;
  float* dbias;
};

template <int VEC, bool FUSE>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy_bf16,
                                                               const float* __restrict__ dy_f32, const float* __restrict__ x,
                                                               const float* __restrict__ w, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ dx_residual, int M, int D,
                                                               float* __restrict__ dx, float* __restrict__ dw,
                                                               float* __restrict__ db, const LnFuse f) {
  // Per-warp column partials of dw / db (/ dbias) live in SHARED memory, not registers: every lane owns its own 16-byte
  // slots (conflict-free), a row costs 2 x NRED x VEC shared accesses per lane, and the kernel drops from ~190 to ~110
  // registers — two blocks (16 warps, 16 rows in flight) per SM instead of one.  The kernel is HBM-bound: rows in flight
  // are what buys bandwidth (r01: 2.7 TB/s with 8 warps per SM).
  extern __shared__ float red[];  // [warps][NRED * D]
  constexpr int NRED = FUSE ? 3 : 2;
  DropCfg dc = make_drop(0.f);
  uint64_t doff = 0;
  if (FUSE) {
    dc = make_drop(f.dropout_p);
    doff = f.offset + ((f.dropout_p > 0.f && f.offset_dev) ? __ldg(f.offset_dev) : 0ull);
  }
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int gw = blockIdx.x * warps_per_block + wib;
  const int total_warps = gridDim.x * warps_per_block;
  const int D4 = D / 4;
  float4* acc4 = reinterpret_cast<float4*>(red) + wib * NRED * D4;  // this warp's [NRED][D4] partials
#pragma unroll
  for (int j = 0; j < NRED * VEC; ++j) acc4[lane + 32 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ww[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) ww[j] = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * j);
  for (int row = gw; row < M; row += total_warps) {
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * D);
    float4 xh[VEC], g[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c4 = lane + 32 * j;
      const float4 xv = xr[c4];
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dy_f32) d = reinterpret_cast<const float4*>(dy_f32 + static_cast<int64_t>(row) * D)[c4];
      if (dy_bf16) {  // both given: dy = dy_f32 + dy_bf16 (residual-stream grad + branch grad)
        const uint2 pk = reinterpret_cast<const uint2*>(dy_bf16 + static_cast<int64_t>(row) * D)[c4];
        d.x += bf16_lo(pk.x); d.y += bf16_hi(pk.x); d.z += bf16_lo(pk.y); d.w += bf16_hi(pk.y);
      }
      xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[j] = make_float4(d.x * ww[j].x, d.y * ww[j].y, d.z * ww[j].z, d.w * ww[j].w);
      s1 += g[j].x + g[j].y + g[j].z + g[j].w;
      s2 += g[j].x * xh[j].x + g[j].y * xh[j].y + g[j].z * xh[j].z + g[j].w * xh[j].w;
      float4 aw = acc4[c4], ab = acc4[D4 + c4];
      aw.x += d.x * xh[j].x; aw.y += d.y * xh[j].y; aw.z += d.z * xh[j].z; aw.w += d.w * xh[j].w;
      ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
      acc4[c4] = aw;
      acc4[D4 + c4] = ab;
    }
    const float c1 = warp_sum(s1) / D, c2 = warp_sum(s2) / D;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c4 = lane + 32 * j;
      float4 o;
      o.x = rs * (g[j].x - c1 - xh[j].x * c2);
      o.y = rs * (g[j].y - c1 - xh[j].y * c2);
      o.z = rs * (g[j].z - c1 - xh[j].z * c2);
      o.w = rs * (g[j].w - c1 - xh[j].w * c2);
      if (dx_residual) {
        const float4 r = reinterpret_cast<const float4*>(dx_residual + static_cast<int64_t>(row) * D)[c4];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      reinterpret_cast<float4*>(dx + static_cast<int64_t>(row) * D)[c4] = o;
      if (FUSE) {
        if (f.dropout_p > 0.f) {
          const uint64_t idx = static_cast<uint64_t>(row) * static_cast<uint64_t>(D) + static_cast<uint64_t>(c4 * 4);
          float k[8];
          drop8(f.seed, doff, idx >> 3, dc, k);
          const int h = static_cast<int>(idx & 4);  // this lane's 4 columns are the low or the high half of the group
          o.x *= k[h]; o.y *= k[h + 1]; o.z *= k[h + 2]; o.w *= k[h + 3];
        }
        float4 ag = acc4[2 * D4 + c4];
        ag.x += o.x; ag.y += o.y; ag.z += o.z; ag.w += o.w;
        acc4[2 * D4 + c4] = ag;
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(f.g + static_cast<int64_t>(row) * D)[c4] = pk;
      }
    }
  }
  // block reduction of the warps' partials: one atomic per column per block
  __syncthreads();
  for (int c = threadIdx.x; c < NRED * D; c += blockDim.x) {
    float s = 0.f;
    for (int ww_ = 0; ww_ < warps_per_block; ++ww_) s += red[ww_ * NRED * D + c];
    if (c < D) atomicAdd(dw + c, s);
    else if (c < 2 * D) atomicAdd(db + (c - D), s);
    else if (f.dbias) atomicAdd(f.dbias + (c - 2 * D), s);
  }
}

// ---------------------------------------------------------------------------------------------
// Column reductions.  Block = 8 warps; a warp covers a contiguous run of columns (lane = 4 fp32 or 8 bf16 columns,
// 512 contiguous bytes per row) and the 8 warps interleave over the block's rows with 4 independent row loads in
// flight per thread; partial sums are combined across the warps in shared memory, so each block issues ONE atomic per
// column (same-address atomics were the bottleneck of the thread-per-column-group version).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scale_cast_colsum_kernel(
    const float* __restrict__ dx, int64_t ld_dx, int M, int N, const float* __restrict__ gamma,
    const float* __restrict__ row_scale, int rows_per_scale, float dropout_p, uint64_t seed, uint64_t offset_host,
    const uint64_t* __restrict__ offset_dev, const __nv_bfloat16* __restrict__ y, int64_t ld_y, __nv_bfloat16* __restrict__ g, int64_t ld_g,
    float* __restrict__ dbias, float* __restrict__ dgamma, int rows_per_block) {
  __shared__ float4 red[2][8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + lane) * 4;
  const bool col_ok = n < N;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
  if (gamma && col_ok) gm = __ldg(reinterpret_cast<const float4*>(gamma + n));
  const DropCfg dc = make_drop(dropout_p);
  const uint64_t offset = offset_host + ((dropout_p > 0.f && offset_dev) ? __ldg(offset_dev) : 0ull);
  float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sg = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col_ok) {
    for (int m0 = r0 + w; m0 < r1; m0 += 32) {  // rows m0, m0+8, m0+16, m0+24: four loads in flight
      float4 d[4];
      uint2 yk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = m0 + 8 * u;
        if (m < r1) {
          d[u] = __ldg(reinterpret_cast<const float4*>(dx + static_cast<int64_t>(m) * ld_dx + n));
          if (dgamma) yk[u] = __ldg(reinterpret_cast<const uint2*>(y + static_cast<int64_t>(m) * ld_y + n));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = m0 + 8 * u;
        if (m >= r1) continue;
        float4 v = d[u];
        if (row_scale) {
          const float sc = __ldg(row_scale + m / rows_per_scale);
          v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        }
        if (dgamma) {
          sg.x += v.x * bf16_lo(yk[u].x); sg.y += v.y * bf16_hi(yk[u].x);
          sg.z += v.z * bf16_lo(yk[u].y); sg.w += v.w * bf16_hi(yk[u].y);
        }
        v.x *= gm.x; v.y *= gm.y; v.z *= gm.z; v.w *= gm.w;
        if (dropout_p > 0.f) {
          const uint64_t idx = static_cast<uint64_t>(m) * static_cast<uint64_t>(N) + static_cast<uint64_t>(n);
          float k[8];
          drop8(seed, offset, idx >> 3, dc, k);
          const int o = static_cast<int>(idx & 4);  // this thread's 4 columns are the low or the high half of the group
          v.x *= k[o]; v.y *= k[o + 1]; v.z *= k[o + 2]; v.w *= k[o + 3];
        }
        sb.x += v.x; sb.y += v.y; sb.z += v.z; sb.w += v.w;
        if (g) {
          uint2 pk;
          pk.x = pack_bf16x2(v.x, v.y);
          pk.y = pack_bf16x2(v.z, v.w);
          *reinterpret_cast<uint2*>(g + static_cast<int64_t>(m) * ld_g + n) = pk;
        }
      }
    }
  }
  red[0][w][lane] = sb;
  red[1][w][lane] = sg;
  __syncthreads();
  if (w < 2 && col_ok) {  // warp 0 finishes dbias, warp 1 dgamma
    float* out = w == 0 ? dbias : dgamma;
    if (out) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 q = red[w][i][lane];
        t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
      }
      atomicAdd(out + n, t.x); atomicAdd(out + n + 1, t.y); atomicAdd(out + n + 2, t.z); atomicAdd(out + n + 3, t.w);
    }
  }
}

__device__ __forceinline__ void acc8(float (&s)[8], const uint4& q) {
  s[0] += bf16_lo(q.x); s[1] += bf16_hi(q.x); s[2] += bf16_lo(q.y); s[3] += bf16_hi(q.y);
  s[4] += bf16_lo(q.z); s[5] += bf16_hi(q.z); s[6] += bf16_lo(q.w); s[7] += bf16_hi(q.w);
}
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int M, int N,
                                                          float* __restrict__ dcol, int rows_per_block) {
  __shared__ float red[8][32][9];  // +1 padding: conflict-free column reads
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + lane) * 8;
  const bool col_ok = n < N;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col_ok) {
    const __nv_bfloat16* px = x + n;
    for (int m0 = r0 + w; m0 < r1; m0 += 32) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (m0 + 8 * u < r1) q[u] = __ldg(reinterpret_cast<const uint4*>(px + static_cast<int64_t>(m0 + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (m0 + 8 * u < r1) acc8(s, q[u]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[w][lane][j] = s[j];
  __syncthreads();
  // 256 threads finish the 32 x 8 columns of the block: thread t -> column (t >> 3, t & 7)
  const int cl = threadIdx.x >> 3, cj = threadIdx.x & 7;
  const int nn = (blockIdx.x * 32 + cl) * 8 + cj;
  if (nn < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cl][cj];
    atomicAdd(dcol + nn, t);
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            int64_t n) {
  const int64_t n8 = n / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 q;
    q.x = pack_bf16x2(a.x, a.y); q.y = pack_bf16x2(a.z, a.w); q.z = pack_bf16x2(b.x, b.y); q.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = q;
  }
  // tail
  for (int64_t i = n8 * 8 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  const int64_t n4 = n / 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int64_t i = n4 * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// Flat AdamW (decoupled weight decay, torch.optim.AdamW semantics):
//   p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g^2;
//   p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         __nv_bfloat16* __restrict__ p_bf16, int64_t n,
                                                         const int64_t* __restrict__ seg_end,
                                                         const float* __restrict__ seg_lr,
                                                         const float* __restrict__ seg_wd, int n_seg, float beta1,
                                                         float beta2, float eps, int step,
                                                         const int32_t* __restrict__ step_dev,
                                                         const float* __restrict__ grad_scale_dev, int zero_grad) {
  const float gs = grad_scale_dev ? *grad_scale_dev : 1.0f;
  const float stepf = static_cast<float>(step_dev ? *step_dev : step);
  const float bc1 = 1.0f - powf(beta1, stepf);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, stepf));
  // 4 consecutive parameters per thread (16-byte accesses).  Arena offsets are multiples of 8 and every parameter
  // that follows another one inside a packed group has a size that is a multiple of 4, so a float4 never straddles
  // two segments: one segment lookup per float4.
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i4 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
    const int64_t i = i4 << 2;
    int lo = 0, hi = n_seg - 1;  // first segment with seg_end > i
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(seg_end + mid) > i) hi = mid; else lo = mid + 1;
    }
    const float lr = __ldg(seg_lr + lo), wd = __ldg(seg_wd + lo);
    const float decay = 1.0f - lr * wd, step_size = lr / bc1;
    float4 pv = reinterpret_cast<float4*>(p)[i4];
    const float4 gv = reinterpret_cast<const float4*>(g)[i4];
    if (zero_grad) reinterpret_cast<float4*>(g)[i4] = make_float4(0.f, 0.f, 0.f, 0.f);  // next step's zero_grad, for 4 B / param
    float4 mv = reinterpret_cast<float4*>(m)[i4];
    float4 vv = reinterpret_cast<float4*>(v)[i4];
#define X2K_ADAM1(c)                                                         \
    {                                                                        \
      const float gi = gv.c * gs;                                            \
      mv.c = beta1 * mv.c + (1.0f - beta1) * gi;                             \
      vv.c = beta2 * vv.c + (1.0f - beta2) * gi * gi;                        \
      pv.c = pv.c * decay - step_size * mv.c / (sqrtf(vv.c) / bc2_sqrt + eps); \
    }
    X2K_ADAM1(x) X2K_ADAM1(y) X2K_ADAM1(z) X2K_ADAM1(w)
#undef X2K_ADAM1
    reinterpret_cast<float4*>(p)[i4] = pv;
    reinterpret_cast<float4*>(m)[i4] = mv;
    reinterpret_cast<float4*>(v)[i4] = vv;
    if (p_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(pv.x, pv.y);
      pk.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(p_bf16)[i4] = pk;
    }
  }
}

// out[s, :] = sum over rows b with index[b] == s of in[b, :]   (fp32 accumulate, bf16 in/out)
__global__ void __launch_bounds__(128) segment_sum_bf16_kernel(const __nv_bfloat16* __restrict__ in,
                                                               const int32_t* __restrict__ index, int n_rows,
                                                               int64_t row_elems, __nv_bfloat16* __restrict__ out) {
  extern __shared__ int32_t members[];  // rows of this segment
  __shared__ int n_members;
  const int seg = blockIdx.y;
  if (threadIdx.x == 0) n_members = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < n_rows; b += blockDim.x)
    if (index[b] == seg) members[atomicAdd(&n_members, 1)] = b;
  __syncthreads();
  const int nm = n_members;
  const int64_t c8 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (c8 >= row_elems) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < nm; ++i) {
    const uint4 q = *reinterpret_cast<const uint4*>(in + static_cast<int64_t>(members[i]) * row_elems + c8);
    acc[0] += bf16_lo(q.x); acc[1] += bf16_hi(q.x); acc[2] += bf16_lo(q.y); acc[3] += bf16_hi(q.y);
    acc[4] += bf16_lo(q.z); acc[5] += bf16_hi(q.z); acc[6] += bf16_lo(q.w); acc[7] += bf16_hi(q.w);
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + static_cast<int64_t>(seg) * row_elems + c8) = o;
}

template <typename F>
int dispatch_vec(int D, F&& f) {
  switch (D / 128) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 3: return f(std::integral_constant<int, 3>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 5: return f(std::integral_constant<int, 5>{});
    case 6: return f(std::integral_constant<int, 6>{});
    case 7: return f(std::integral_constant<int, 7>{});
    case 8: return f(std::integral_constant<int, 8>{});
  }
  return X2K_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace x2k

using namespace x2k;

extern "C" int x2k_layernorm_fwd(const float* x, const float* w, const float* b, int32_t M, int32_t D, float eps,
                                 void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(x && w && b && (y_bf16 || y_f32), "x2k_layernorm_fwd: NULL argument");
  X2K_REQUIRE(M > 0 && D > 0 && D % 128 == 0 && D <= 1024, "x2k_layernorm_fwd: D=%d must be a multiple of 128, <= 1024", D);
  const int rows_per_block = 8;
  const int grid = (M + rows_per_block - 1) / rows_per_block;
  return dispatch_vec(D, [&](auto vec) {
    layernorm_fwd_kernel<decltype(vec)::value><<<grid, 256, 0, stream>>>(
        x, w, b, M, D, eps, static_cast<__nv_bfloat16*>(y_bf16), y_f32, mean, rstd);
    X2K_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return X2K_OK;
  });
}

extern "C" int x2k_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* w,
                                 const float* mean, const float* rstd, const float* dx_residual, int32_t M, int32_t D,
                                 float* dx, float* dw, float* db, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(dy_bf16 != nullptr || dy_f32 != nullptr, "x2k_layernorm_bwd: dy_bf16 and/or dy_f32 must be given");
  X2K_REQUIRE(x && w && mean && rstd && dx && dw && db, "x2k_layernorm_bwd: NULL argument");
  X2K_REQUIRE(M > 0 && D % 128 == 0 && D <= 1024, "x2k_layernorm_bwd: D=%d must be a multiple of 128, <= 1024", D);
  const int warps = 8;
  int grid = sm_count() * 4;  // two resident blocks per SM, two rounds
  if (grid * warps > M) grid = (M + warps - 1) / warps;
  const size_t smem = static_cast<size_t>(warps) * 2 * D * sizeof(float);
  return dispatch_vec(D, [&](auto vec) {
    auto kern = layernorm_bwd_kernel<decltype(vec)::value, false>;
    if (smem > 48 * 1024) X2K_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, warps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, w, mean, rstd,
                                             dx_residual, M, D, dx, dw, db, LnFuse{});
    X2K_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return X2K_OK;
  });
}

extern "C" int x2k_layernorm_bwd_dropcast(const void* dy_bf16, const float* dy_f32, const float* x, const float* w,
                                          const float* mean, const float* rstd, const float* dx_residual, int32_t M,
                                          int32_t D, float* dx, float* dw, float* db, float dropout_p,
                                          uint64_t dropout_seed, uint64_t dropout_offset,
                                          const uint64_t* dropout_offset_dev, void* g_bf16, float* dbias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(dy_bf16 != nullptr || dy_f32 != nullptr, "x2k_layernorm_bwd_dropcast: dy_bf16 and/or dy_f32 must be given");
  X2K_REQUIRE(x && w && mean && rstd && dx && dw && db && g_bf16, "x2k_layernorm_bwd_dropcast: NULL argument");
  X2K_REQUIRE(M > 0 && D % 128 == 0 && D <= 1024, "x2k_layernorm_bwd_dropcast: D=%d must be a multiple of 128, <= 1024", D);
  X2K_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "x2k_layernorm_bwd_dropcast: dropout_p");
  const int warps = 8;
  int grid = sm_count() * 4;  // two resident blocks per SM, two rounds
  if (grid * warps > M) grid = (M + warps - 1) / warps;
  const size_t smem = static_cast<size_t>(warps) * 3 * D * sizeof(float);
  LnFuse f;
  f.dropout_p = dropout_p; f.seed = dropout_seed; f.offset = dropout_offset; f.offset_dev = dropout_offset_dev;
  f.g = static_cast<__nv_bfloat16*>(g_bf16); f.dbias = dbias;
  return dispatch_vec(D, [&](auto vec) {
    auto kern = layernorm_bwd_kernel<decltype(vec)::value, true>;
    if (smem > 48 * 1024) X2K_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, warps * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(dy_bf16), dy_f32, x, w, mean, rstd,
                                             dx_residual, M, D, dx, dw, db, f);
    X2K_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return X2K_OK;
  });
}

extern "C" int x2k_scale_cast_colsum(const float* dx, int64_t ld_dx, int32_t M, int32_t N, const float* gamma,
                                     const float* row_scale, int32_t rows_per_scale, float dropout_p,
                                     uint64_t dropout_seed, uint64_t dropout_offset, const uint64_t* dropout_offset_dev,
                                     const void* y_bf16, int64_t ld_y, void* g_bf16, int64_t ld_g, float* dbias,
                                     float* dgamma, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(dx && M > 0 && N > 0 && N % 4 == 0 && ld_dx % 4 == 0, "x2k_scale_cast_colsum: bad dx/shape");
  X2K_REQUIRE(!dgamma || (y_bf16 && ld_y % 4 == 0), "x2k_scale_cast_colsum: dgamma needs y");
  X2K_REQUIRE(!g_bf16 || ld_g % 4 == 0, "x2k_scale_cast_colsum: ld_g must be a multiple of 4");
  X2K_REQUIRE(!row_scale || rows_per_scale > 0, "x2k_scale_cast_colsum: rows_per_scale");
  X2K_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "x2k_scale_cast_colsum: dropout_p");
  const int col_blocks = (N / 4 + 31) / 32;  // 128 columns per block
  int row_blocks = (sm_count() * 4 + col_blocks - 1) / col_blocks;
  int rows_per_block = (M + row_blocks - 1) / row_blocks;
  if (rows_per_block < 64) rows_per_block = 64;
  row_blocks = (M + rows_per_block - 1) / rows_per_block;
  scale_cast_colsum_kernel<<<dim3(col_blocks, row_blocks), 256, 0, stream>>>(
      dx, ld_dx, M, N, gamma, row_scale, rows_per_scale, dropout_p, dropout_seed, dropout_offset, dropout_offset_dev,
      static_cast<const __nv_bfloat16*>(y_bf16), ld_y, static_cast<__nv_bfloat16*>(g_bf16), ld_g, dbias, dgamma,
      rows_per_block);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int x2k_colsum_bf16(const void* x_bf16, int64_t ld, int32_t M, int32_t N, float* dcol, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(x_bf16 && dcol && M > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "x2k_colsum_bf16: bad arguments");
  const int col_blocks = (N / 8 + 31) / 32;  // 256 columns per block
  int row_blocks = (sm_count() * 4 + col_blocks - 1) / col_blocks;
  int rows_per_block = (M + row_blocks - 1) / row_blocks;
  if (rows_per_block < 64) rows_per_block = 64;
  row_blocks = (M + rows_per_block - 1) / rows_per_block;
  colsum_bf16_kernel<<<dim3(col_blocks, row_blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x_bf16), ld, M,
                                                                       N, dcol, rows_per_block);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int x2k_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(src && dst_bf16 && n > 0, "x2k_cast_f32_bf16: bad arguments");
  X2K_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0,
              "x2k_cast_f32_bf16: pointers must be 16-byte aligned");
  int64_t blocks = (n / 8 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst_bf16), n);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int x2k_sumsq(const float* g, int64_t n, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(g && out && n > 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "x2k_sumsq: bad arguments");
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(g, n, out);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int x2k_adamw_flat(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n,
                              const int64_t* seg_end, const float* seg_lr, const float* seg_wd, int32_t n_seg,
                              float beta1, float beta2, float eps, int32_t step, const int32_t* step_dev,
                              const float* grad_scale_dev, int32_t zero_grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(p && g && m && v && n > 0 && seg_end && seg_lr && seg_wd && n_seg > 0 && (step > 0 || step_dev),
              "x2k_adamw_flat: bad arguments");
  X2K_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(m) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
              "x2k_adamw_flat: buffers must be 16-byte aligned and n a multiple of 4");
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  adamw_flat_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16), n,
                                                                   seg_end, seg_lr, seg_wd, n_seg, beta1, beta2, eps,
                                                                   step, step_dev, grad_scale_dev, zero_grad);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

extern "C" int x2k_segment_sum_bf16(const void* in_bf16, const int32_t* index, int32_t n_rows, int64_t row_elems,
                                    int32_t n_seg, void* out_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(in_bf16 && index && out_bf16 && n_rows > 0 && n_seg > 0 && row_elems > 0 && row_elems % 8 == 0,
              "x2k_segment_sum_bf16: bad arguments");
  X2K_REQUIRE(static_cast<size_t>(n_rows) * sizeof(int32_t) <= 48 * 1024, "x2k_segment_sum_bf16: n_rows too large");
  dim3 grid(static_cast<unsigned>((row_elems / 8 + 127) / 128), n_seg);
  segment_sum_bf16_kernel<<<grid, 128, n_rows * sizeof(int32_t), stream>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), index, n_rows, row_elems, static_cast<__nv_bfloat16*>(out_bf16));
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}

// ---------------------------------------------------------------------------------------------
// fused cross entropy, second half: reduce the per-16-column online-softmax partials of x2k_gemm(ce_mode 1)
// ---------------------------------------------------------------------------------------------
namespace x2k {
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float2* __restrict__ partials, const float* __restrict__ tlogit, const int64_t* __restrict__ labels,
                   int M, int groups, float* __restrict__ lse, float* __restrict__ loss) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float2* pr = partials + static_cast<int64_t>(row) * groups;
  float mx = -INFINITY, sum = 0.f;
  for (int g = lane; g < groups; g += 32) {
    const float2 v = __ldg(pr + g);
    const float nm = fmaxf(mx, v.x);
    if (nm > -INFINITY) sum = sum * __expf(mx - nm) + v.y * __expf(v.x - nm);
    mx = nm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
    const float nm = fmaxf(mx, om);
    if (nm > -INFINITY) sum = sum * __expf(mx - nm) + os * __expf(om - nm);
    mx = nm;
  }
  if (lane == 0) {
    const float l = mx + __logf(sum);
    lse[row] = l;
    const long long lbl = labels[row];
    loss[row] = lbl >= 0 ? l - tlogit[row] : 0.f;
  }
}
}  // namespace x2k

extern "C" int x2k_ce_finalize(const float* partials, const float* target_logit, const int64_t* labels, int32_t M, int32_t N,
                               float* lse, float* loss, void* stream_) {
  using namespace x2k;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  X2K_REQUIRE(partials && target_logit && labels && lse && loss && M > 0 && N > 0, "x2k_ce_finalize: bad arguments");
  X2K_REQUIRE((reinterpret_cast<uintptr_t>(partials) & 7) == 0, "x2k_ce_finalize: partials must be 8-byte aligned");
  ce_finalize_kernel<<<(M + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const float2*>(partials), target_logit, labels, M,
                                                      (N + 15) / 16, lse, loss);
  X2K_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return X2K_OK;
}
