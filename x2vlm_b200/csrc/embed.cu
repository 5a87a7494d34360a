// embed.cu — the two small HBM-bound ends of the encoders, sm_100a.
//
//   x2k_embed_ln_{fwd,bwd}   BertEmbeddings (models/xbert.py:189-216): y = dropout(LayerNorm(word[id] + pos[p] + type[t])).
//                            One warp per token row: three 16-byte-vectorised gathers, LayerNorm statistics by warp
//                            shuffles, Philox dropout, one coalesced fp32 (+ bf16) store — instead of three gather
//                            kernels, two adds, a LayerNorm and a dropout kernel.  Backward recomputes the summed row
//                            (nothing but mean / rstd is saved), applies dropout' and LayerNorm', and scatters the row
//                            gradient into the three tables with 16-byte vector reductions (red.global.add.v4.f32).
//   x2k_pool_tail_{fwd,bwd}  the tail of VisionTransformer.forward (models/beit2.py:409-436): drop the cls output,
//                            fc_norm over the patch tokens, their mean as the new token 0 — and in region mode the
//                            per-region gather + mask-weighted mean (image_atts[:, 1:]).  One block per output
//                            sequence; rows normalised by warps, the (weighted) mean accumulated in shared memory.
#include <type_traits>

#include "common.cuh"

namespace x2k {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct EmbedParams {
  const int64_t* ids;       // [M]
  const int64_t* type_ids;  // [M] or nullptr (= 0)
  const int64_t* pos_ids;   // [M] or nullptr (= pos_offset + m % L)
  int M, D, L, pos_offset;
  const float *word, *pos, *type;  // [V,D], [P,D], [T,D]
  const float *ln_w, *ln_b;
  float eps, dropout_p;
  uint64_t seed, offset;
  const uint64_t* offset_dev;
};

template <int VEC>
__device__ __forceinline__ void gather_row(const EmbedParams& p, int row, int lane, float4 (&v)[VEC], int64_t& id, int64_t& pp, int64_t& tt) {
  id = p.ids[row];
  tt = p.type_ids ? p.type_ids[row] : 0;
  pp = p.pos_ids ? p.pos_ids[row] : static_cast<int64_t>(p.pos_offset + row % p.L);
  const float4* w = reinterpret_cast<const float4*>(p.word + id * p.D);
  const float4* q = reinterpret_cast<const float4*>(p.pos + pp * p.D);
  const float4* t = reinterpret_cast<const float4*>(p.type + tt * p.D);
#pragma unroll
  for (int j = 0; j < VEC; ++j) v[j] = add4(add4(__ldg(w + lane + 32 * j), __ldg(q + lane + 32 * j)), __ldg(t + lane + 32 * j));
}

template <int VEC>
__global__ void __launch_bounds__(256)
embed_ln_fwd_kernel(const EmbedParams p, float* __restrict__ y_f32, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ mean_out,
                    float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.M) return;
  float4 v[VEC];
  int64_t id, pp, tt;
  gather_row<VEC>(p, row, lane, v, id, pp, tt);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) s += v[j].x + v[j].y + v[j].z + v[j].w;
  const float mean = warp_sum(s) / p.D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
    q += a * a + c * c + d * d + e * e;
  }
  const float rstd = rsqrtf(warp_sum(q) / p.D + p.eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const DropCfg dc = make_drop(p.dropout_p);
  const uint64_t doff = p.offset + ((p.dropout_p > 0.f && p.offset_dev) ? __ldg(p.offset_dev) : 0ull);
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int c4 = lane + 32 * j;
    const float4 ww = __ldg(reinterpret_cast<const float4*>(p.ln_w) + c4);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.ln_b) + c4);
    float4 o;
    o.x = (v[j].x - mean) * rstd * ww.x + bb.x;
    o.y = (v[j].y - mean) * rstd * ww.y + bb.y;
    o.z = (v[j].z - mean) * rstd * ww.z + bb.z;
    o.w = (v[j].w - mean) * rstd * ww.w + bb.w;
    if (p.dropout_p > 0.f) {
      const uint64_t idx = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.D) + static_cast<uint64_t>(c4 * 4);
      float k[8];
      drop8(p.seed, doff, idx >> 3, dc, k);
      const int h = static_cast<int>(idx & 4);
      o.x *= k[h]; o.y *= k[h + 1]; o.z *= k[h + 2]; o.w *= k[h + 3];
    }
    reinterpret_cast<float4*>(y_f32 + static_cast<int64_t>(row) * p.D)[c4] = o;
    if (y_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y_bf16 + static_cast<int64_t>(row) * p.D)[c4] = pk;
    }
  }
}

// Backward: warps stride over the rows.  LayerNorm weight / bias gradients are kept per column in registers and reduced
// through shared memory (one atomic per column per block); the row gradient goes to the three tables by vector reductions.
template <int VEC>
__global__ void __launch_bounds__(256)
embed_ln_bwd_kernel(const EmbedParams p, const float* __restrict__ dy, const float* __restrict__ mean, const float* __restrict__ rstd,
                    float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtype, float* __restrict__ dw,
                    float* __restrict__ db) {
  extern __shared__ float red[];  // [warps][2 * D]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int gw = blockIdx.x * warps_per_block + wib, total_warps = gridDim.x * warps_per_block;
  const DropCfg dc = make_drop(p.dropout_p);
  const uint64_t doff = p.offset + ((p.dropout_p > 0.f && p.offset_dev) ? __ldg(p.offset_dev) : 0ull);
  float4 ww[VEC], adw[VEC], adb[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    ww[j] = __ldg(reinterpret_cast<const float4*>(p.ln_w) + lane + 32 * j);
    adw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    adb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = gw; row < p.M; row += total_warps) {
    float4 xh[VEC], g[VEC];
    int64_t id, pp, tt;
    gather_row<VEC>(p, row, lane, xh, id, pp, tt);
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c4 = lane + 32 * j;
      float4 d = reinterpret_cast<const float4*>(dy + static_cast<int64_t>(row) * p.D)[c4];
      if (p.dropout_p > 0.f) {
        const uint64_t idx = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.D) + static_cast<uint64_t>(c4 * 4);
        float k[8];
        drop8(p.seed, doff, idx >> 3, dc, k);
        const int h = static_cast<int>(idx & 4);
        d.x *= k[h]; d.y *= k[h + 1]; d.z *= k[h + 2]; d.w *= k[h + 3];
      }
      xh[j] = make_float4((xh[j].x - mu) * rs, (xh[j].y - mu) * rs, (xh[j].z - mu) * rs, (xh[j].w - mu) * rs);
      adb[j] = add4(adb[j], d);
      adw[j].x += d.x * xh[j].x; adw[j].y += d.y * xh[j].y; adw[j].z += d.z * xh[j].z; adw[j].w += d.w * xh[j].w;
      g[j] = make_float4(d.x * ww[j].x, d.y * ww[j].y, d.z * ww[j].z, d.w * ww[j].w);
      s1 += g[j].x + g[j].y + g[j].z + g[j].w;
      s2 += g[j].x * xh[j].x + g[j].y * xh[j].y + g[j].z * xh[j].z + g[j].w * xh[j].w;
    }
    const float c1 = warp_sum(s1) / p.D, c2 = warp_sum(s2) / p.D;
    float* rw = dword + id * p.D;
    float* rp = dpos + pp * p.D;
    float* rt = dtype + tt * p.D;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c4 = lane + 32 * j;
      float4 o;
      o.x = rs * (g[j].x - c1 - xh[j].x * c2);
      o.y = rs * (g[j].y - c1 - xh[j].y * c2);
      o.z = rs * (g[j].z - c1 - xh[j].z * c2);
      o.w = rs * (g[j].w - c1 - xh[j].w * c2);
      red_add_v4(rw + c4 * 4, o);
      red_add_v4(rp + c4 * 4, o);
      red_add_v4(rt + c4 * 4, o);
    }
  }
  float4* red4 = reinterpret_cast<float4*>(red);
  const int D4 = p.D / 4;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    red4[wib * 2 * D4 + lane + 32 * j] = adw[j];
    red4[wib * 2 * D4 + D4 + lane + 32 * j] = adb[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * p.D; c += blockDim.x) {
    float s = 0.f;
    for (int w_ = 0; w_ < warps_per_block; ++w_) s += red[w_ * 2 * p.D + c];
    if (c < p.D) atomicAdd(dw + c, s);
    else atomicAdd(db + (c - p.D), s);
  }
}

// ---------------------------------------------------------------------------------------------
// vision tail
// ---------------------------------------------------------------------------------------------
struct PoolParams {
  const float* x;            // [n_img, N, D] block output (token 0 = cls, dropped)
  int n_img, n_out, N, D;
  const int64_t* group;      // [n_out] image of every output sequence, or nullptr (identity)
  const int64_t* atts;       // [n_out, N] 0/1 region masks (column 0 = cls, ignored), or nullptr (plain mean)
  const float *w, *b;
  float eps;
};

// fwd: block = one output sequence; its 8 warps walk the N-1 patch rows: LayerNorm -> out[:, 1 + r]; the (mask-weighted)
// sum of the normalised rows is accumulated per warp in registers, combined in shared memory -> out[:, 0] = sum / weight.
template <int VEC>
__global__ void __launch_bounds__(256)
pool_tail_fwd_kernel(const PoolParams p, float* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  extern __shared__ float red[];  // [8][D]
  const int s = blockIdx.x, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int img = p.group ? static_cast<int>(p.group[s]) : s;
  const float* xi = p.x + static_cast<int64_t>(img) * p.N * p.D;
  float* os = out + static_cast<int64_t>(s) * p.N * p.D;
  float4 acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float wsum = 0.f;
  for (int r = 1 + wib; r < p.N; r += 8) {
    const float4* xr = reinterpret_cast<const float4*>(xi + static_cast<int64_t>(r) * p.D);
    float4 v[VEC];
    float sm = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      v[j] = xr[lane + 32 * j];
      sm += v[j].x + v[j].y + v[j].z + v[j].w;
    }
    const float mean = warp_sum(sm) / p.D;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
      q += a * a + c * c + d * d + e * e;
    }
    const float rstd = rsqrtf(warp_sum(q) / p.D + p.eps);
    if (lane == 0 && mean_out) {  // statistics per (output sequence, patch row): the backward re-normalises with them
      mean_out[static_cast<int64_t>(s) * p.N + r] = mean;
      rstd_out[static_cast<int64_t>(s) * p.N + r] = rstd;
    }
    const float wt = p.atts ? static_cast<float>(p.atts[static_cast<int64_t>(s) * p.N + r]) : 1.0f;
    wsum += wt;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c4 = lane + 32 * j;
      const float4 ww = __ldg(reinterpret_cast<const float4*>(p.w) + c4);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b) + c4);
      float4 o;
      o.x = (v[j].x - mean) * rstd * ww.x + bb.x;
      o.y = (v[j].y - mean) * rstd * ww.y + bb.y;
      o.z = (v[j].z - mean) * rstd * ww.z + bb.z;
      o.w = (v[j].w - mean) * rstd * ww.w + bb.w;
      reinterpret_cast<float4*>(os + static_cast<int64_t>(r) * p.D)[c4] = o;
      acc[j].x += wt * o.x; acc[j].y += wt * o.y; acc[j].z += wt * o.z; acc[j].w += wt * o.w;
    }
  }
  float4* red4 = reinterpret_cast<float4*>(red);
  const int D4 = p.D / 4;
#pragma unroll
  for (int j = 0; j < VEC; ++j) red4[wib * D4 + lane + 32 * j] = acc[j];
  __shared__ float wred[8];
  if (lane == 0) wred[wib] = wsum;
  __syncthreads();
  float wt = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) wt += wred[i];
  for (int c = threadIdx.x; c < p.D; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i * p.D + c];
    os[c] = t / wt;  // token 0: (weighted) mean of the normalised patch tokens
  }
}

// bwd: dx[img, r] += LN'( d_out[s, r] + wt[s, r] / W[s] * d_out[s, 0] ) for every output sequence s of image img
// (vector reductions: several region samples share one image); dw / db of fc_norm through the block reduction;
// the cls row of dx receives nothing (the cls output is dropped).
template <int VEC>
__global__ void __launch_bounds__(256)
pool_tail_bwd_kernel(const PoolParams p, const float* __restrict__ d_out, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float red[];  // [8][2 * D]
  const int s = blockIdx.x, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int img = p.group ? static_cast<int>(p.group[s]) : s;
  const float* xi = p.x + static_cast<int64_t>(img) * p.N * p.D;
  const float* ds = d_out + static_cast<int64_t>(s) * p.N * p.D;
  float* dxi = dx + static_cast<int64_t>(img) * p.N * p.D;
  __shared__ float wtot_s;
  if (threadIdx.x < 32) {
    float w = 0.f;
    for (int r = 1 + lane; r < p.N; r += 32) w += p.atts ? static_cast<float>(p.atts[static_cast<int64_t>(s) * p.N + r]) : 1.0f;
    w = warp_sum(w);
    if (lane == 0) wtot_s = w;
  }
  __syncthreads();
  const float inv_w = 1.0f / wtot_s;
  float4 ww[VEC], d0[VEC], adw[VEC], adb[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    ww[j] = __ldg(reinterpret_cast<const float4*>(p.w) + lane + 32 * j);
    d0[j] = reinterpret_cast<const float4*>(ds)[lane + 32 * j];  // gradient of token 0 (the pooled mean)
    adw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    adb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = 1 + wib; r < p.N; r += 8) {
    const float mu = mean_in[static_cast<int64_t>(s) * p.N + r], rs = rstd_in[static_cast<int64_t>(s) * p.N + r];
    const float wt = (p.atts ? static_cast<float>(p.atts[static_cast<int64_t>(s) * p.N + r]) : 1.0f) * inv_w;
    const float4* xr = reinterpret_cast<const float4*>(xi + static_cast<int64_t>(r) * p.D);
    float4 xh[VEC], g[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c4 = lane + 32 * j;
      const float4 xv = xr[c4];
      float4 d = reinterpret_cast<const float4*>(ds + static_cast<int64_t>(r) * p.D)[c4];
      d.x += wt * d0[j].x; d.y += wt * d0[j].y; d.z += wt * d0[j].z; d.w += wt * d0[j].w;
      xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      adb[j] = add4(adb[j], d);
      adw[j].x += d.x * xh[j].x; adw[j].y += d.y * xh[j].y; adw[j].z += d.z * xh[j].z; adw[j].w += d.w * xh[j].w;
      g[j] = make_float4(d.x * ww[j].x, d.y * ww[j].y, d.z * ww[j].z, d.w * ww[j].w);
      s1 += g[j].x + g[j].y + g[j].z + g[j].w;
      s2 += g[j].x * xh[j].x + g[j].y * xh[j].y + g[j].z * xh[j].z + g[j].w * xh[j].w;
    }
    const float c1 = warp_sum(s1) / p.D, c2 = warp_sum(s2) / p.D;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c4 = lane + 32 * j;
      float4 o;
      o.x = rs * (g[j].x - c1 - xh[j].x * c2);
      o.y = rs * (g[j].y - c1 - xh[j].y * c2);
      o.z = rs * (g[j].z - c1 - xh[j].z * c2);
      o.w = rs * (g[j].w - c1 - xh[j].w * c2);
      // vector reduction: several region samples may point at the same image, and the caller may accumulate the
      // gradients of the full-image and the region outputs into one dx
      red_add_v4(dxi + static_cast<int64_t>(r) * p.D + c4 * 4, o);
    }
  }
  float4* red4 = reinterpret_cast<float4*>(red);
  const int D4 = p.D / 4;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    red4[wib * 2 * D4 + lane + 32 * j] = adw[j];
    red4[wib * 2 * D4 + D4 + lane + 32 * j] = adb[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * p.D; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i * 2 * p.D + c];
    if (c < p.D) atomicAdd(dw + c, t);
    else atomicAdd(db + (c - p.D), t);
  }
}

template <typename F>
int dispatch_vec(int D, const char* who, F&& f) {
  X2K_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, "%s: D = %d must be a multiple of 128 in [128, 1024]", who, D);
  switch (D / 128) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 5: return f(std::integral_constant<int, 5>());
    case 6: return f(std::integral_constant<int, 6>());
    case 7: return f(std::integral_constant<int, 7>());
    default: return f(std::integral_constant<int, 8>());
  }
}

}  // namespace
}  // namespace x2k

using namespace x2k;

static int fill_embed(EmbedParams& p, const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int32_t M, int32_t D,
                      int32_t L, int32_t pos_offset, const float* word, const float* pos, const float* type, const float* ln_w,
                      const float* ln_b, float eps, float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                      const char* who) {
  X2K_REQUIRE(ids && word && pos && type && ln_w && ln_b && M > 0 && L > 0, "%s: bad arguments", who);
  X2K_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "%s: dropout_p", who);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  X2K_REQUIRE(al16(word) && al16(pos) && al16(type) && al16(ln_w) && al16(ln_b), "%s: tables must be 16-byte aligned", who);
  p.ids = ids; p.type_ids = type_ids; p.pos_ids = pos_ids;
  p.M = M; p.D = D; p.L = L; p.pos_offset = pos_offset;
  p.word = word; p.pos = pos; p.type = type; p.ln_w = ln_w; p.ln_b = ln_b;
  p.eps = eps; p.dropout_p = dropout_p; p.seed = seed; p.offset = offset; p.offset_dev = offset_dev;
  return X2K_OK;
}

extern "C" int x2k_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int32_t M, int32_t D, int32_t L,
                                int32_t pos_offset, const float* word, const float* pos, const float* type, const float* ln_w,
                                const float* ln_b, float eps, float dropout_p, uint64_t seed, uint64_t offset,
                                const uint64_t* offset_dev, float* y_f32, void* y_bf16, float* mean, float* rstd, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EmbedParams p;
  if (int rc = fill_embed(p, ids, type_ids, pos_ids, M, D, L, pos_offset, word, pos, type, ln_w, ln_b, eps, dropout_p, seed, offset,
                          offset_dev, "x2k_embed_ln_fwd")) return rc;
  X2K_REQUIRE(y_f32 && mean && rstd, "x2k_embed_ln_fwd: y_f32 / mean / rstd are required");
  return dispatch_vec(D, "x2k_embed_ln_fwd", [&](auto vec) {
    constexpr int VEC = decltype(vec)::value;
    embed_ln_fwd_kernel<VEC><<<(M + 7) / 8, 256, 0, stream>>>(p, y_f32, static_cast<__nv_bfloat16*>(y_bf16), mean, rstd);
    X2K_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return X2K_OK;
  });
}

extern "C" int x2k_embed_ln_bwd(const float* dy, const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int32_t M,
                                int32_t D, int32_t L, int32_t pos_offset, const float* word, const float* pos, const float* type,
                                const float* ln_w, const float* ln_b, const float* mean, const float* rstd, float dropout_p,
                                uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float* dword, float* dpos, float* dtype,
                                float* dw, float* db, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EmbedParams p;
  if (int rc = fill_embed(p, ids, type_ids, pos_ids, M, D, L, pos_offset, word, pos, type, ln_w, ln_b, 0.f, dropout_p, seed, offset,
                          offset_dev, "x2k_embed_ln_bwd")) return rc;
  X2K_REQUIRE(dy && mean && rstd && dword && dpos && dtype && dw && db, "x2k_embed_ln_bwd: NULL argument");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  X2K_REQUIRE(al16(dy) && al16(dword) && al16(dpos) && al16(dtype), "x2k_embed_ln_bwd: 16-byte alignment");
  return dispatch_vec(D, "x2k_embed_ln_bwd", [&](auto vec) {
    constexpr int VEC = decltype(vec)::value;
    int blocks = (M + 63) / 64;  // 8 rows per warp
    if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
    if (blocks < 1) blocks = 1;
    const size_t smem = 8 * 2 * static_cast<size_t>(D) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      X2K_CHECK_CUDA(cudaFuncSetAttribute(embed_ln_bwd_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      attr_set = true;
    }
    embed_ln_bwd_kernel<VEC><<<blocks, 256, smem, stream>>>(p, dy, mean, rstd, dword, dpos, dtype, dw, db);
    X2K_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return X2K_OK;
  });
}

static int fill_pool(PoolParams& p, const float* x, int32_t n_img, int32_t n_out, int32_t N, int32_t D, const int64_t* group,
                     const int64_t* atts, const float* w, const float* b, float eps, const char* who) {
  X2K_REQUIRE(x && w && b && n_img > 0 && n_out > 0 && N > 1, "%s: bad arguments", who);
  X2K_REQUIRE(group != nullptr || n_out == n_img, "%s: without a group index n_out must equal n_img", who);
  X2K_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "%s: x must be 16-byte aligned", who);
  p.x = x; p.n_img = n_img; p.n_out = n_out; p.N = N; p.D = D; p.group = group; p.atts = atts; p.w = w; p.b = b; p.eps = eps;
  return X2K_OK;
}

extern "C" int x2k_pool_tail_fwd(const float* x, int32_t n_img, int32_t n_out, int32_t N, int32_t D, const int64_t* group,
                                 const int64_t* atts, const float* w, const float* b, float eps, float* out, float* mean,
                                 float* rstd, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PoolParams p;
  if (int rc = fill_pool(p, x, n_img, n_out, N, D, group, atts, w, b, eps, "x2k_pool_tail_fwd")) return rc;
  X2K_REQUIRE(out && ((mean == nullptr) == (rstd == nullptr)), "x2k_pool_tail_fwd: out is required; mean / rstd come together");
  return dispatch_vec(D, "x2k_pool_tail_fwd", [&](auto vec) {
    constexpr int VEC = decltype(vec)::value;
    const size_t smem = 8 * static_cast<size_t>(D) * sizeof(float);
    pool_tail_fwd_kernel<VEC><<<n_out, 256, smem, stream>>>(p, out, mean, rstd);
    X2K_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return X2K_OK;
  });
}

extern "C" int x2k_pool_tail_bwd(const float* d_out, const float* x, int32_t n_img, int32_t n_out, int32_t N, int32_t D,
                                 const int64_t* group, const int64_t* atts, const float* w, const float* mean, const float* rstd,
                                 float* dx, int32_t accumulate, float* dw, float* db, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PoolParams p;
  if (int rc = fill_pool(p, x, n_img, n_out, N, D, group, atts, w, w, 0.f, "x2k_pool_tail_bwd")) return rc;
  X2K_REQUIRE(d_out && mean && rstd && dx && dw && db, "x2k_pool_tail_bwd: NULL argument");
  return dispatch_vec(D, "x2k_pool_tail_bwd", [&](auto vec) {
    constexpr int VEC = decltype(vec)::value;
    // dx is reduced into (the cls rows receive nothing): start from zero unless the caller accumulates
    if (!accumulate) X2K_CHECK_CUDA(cudaMemsetAsync(dx, 0, static_cast<size_t>(n_img) * N * D * sizeof(float), stream));
    const size_t smem = 8 * 2 * static_cast<size_t>(D) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      X2K_CHECK_CUDA(cudaFuncSetAttribute(pool_tail_bwd_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      attr_set = true;
    }
    pool_tail_bwd_kernel<VEC><<<n_out, 256, smem, stream>>>(p, d_out, mean, rstd, dx, dw, db);
    X2K_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return X2K_OK;
  });
}
