// attn_common.cuh — pieces shared by the attention kernels (attn.cu: whole key range in TMEM, Lk <= 256;
// attn_long.cu: key-blocked online-softmax kernels for any length): parameter block, 128B-swizzle addressing,
// coalesced bias staging, argument checks.
#pragma once
#include "common.cuh"

namespace x2k {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  int B, H, Lq, Lk, Lk_pad;
  const int32_t* kv_index;
  float scale_log2;  // scale * log2(e)
  float scale;
  const float* bias;
  int64_t bias_h_stride, bias_q_stride;
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
  __nv_bfloat16* ds_out;
  int64_t ds_b_stride, ds_h_stride, ds_q_stride;
  const float* delta;  // [B,H,Lq] rowsum(dO ∘ O) from attn_delta_kernel, or nullptr (computed in the kernel)
};

// byte offset of the 16-byte chunk holding elements [k0, k0+8) of row `row` inside a K-major,
// 128B-swizzled tile made of [128 rows x 64 elements] blocks (block stride 16 KB).
__device__ __forceinline__ uint32_t swz_off(int row, int k0) {
  const int blk = k0 >> 6, c = (k0 & 63) >> 3;
  return blk * 16384 + row * 128 + ((c ^ (row & 7)) << 4);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// additive term (bias + mask) * log2e for 16 consecutive keys starting at k0 of query row q.
__device__ __forceinline__ void load_additive(const AttnParams& p, const float* bias_row, const float* mask_row, int k0,
                                              float (&add)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) add[j] = 0.f;
  if (bias_row) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias_row + k0 + j));
      add[j] += b.x; add[j + 1] += b.y; add[j + 2] += b.z; add[j + 3] += b.w;
    }
  }
  if (mask_row) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mask_row + k0 + j));
      add[j] += m.x; add[j + 1] += m.y; add[j + 2] += m.z; add[j + 3] += m.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) add[j] *= kLog2e;
}

// TMEM row (this thread's lane) -> NCH x 16 fp32 columns -> scaled bf16 -> global (16-byte stores)
template <int NCH>
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, uint32_t tcol_addr, float mul, bool valid) {
  uint32_t o[16];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    tmem_ld_32x16(tcol_addr + c * 16, o);
    tmem_wait_ld();
    if (valid) {
      uint4 v0, v1;
      v0.x = pack_bf16x2(__uint_as_float(o[0]) * mul, __uint_as_float(o[1]) * mul);
      v0.y = pack_bf16x2(__uint_as_float(o[2]) * mul, __uint_as_float(o[3]) * mul);
      v0.z = pack_bf16x2(__uint_as_float(o[4]) * mul, __uint_as_float(o[5]) * mul);
      v0.w = pack_bf16x2(__uint_as_float(o[6]) * mul, __uint_as_float(o[7]) * mul);
      v1.x = pack_bf16x2(__uint_as_float(o[8]) * mul, __uint_as_float(o[9]) * mul);
      v1.y = pack_bf16x2(__uint_as_float(o[10]) * mul, __uint_as_float(o[11]) * mul);
      v1.z = pack_bf16x2(__uint_as_float(o[12]) * mul, __uint_as_float(o[13]) * mul);
      v1.w = pack_bf16x2(__uint_as_float(o[14]) * mul, __uint_as_float(o[15]) * mul);
      *reinterpret_cast<uint4*>(dst + c * 16) = v0;
      *reinterpret_cast<uint4*>(dst + c * 16 + 8) = v1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Thread mapping shared by both kernels: 256 threads = 8 warps.  Warp w owns TMEM lane quadrant (w & 3) — the
// hardware restriction for tcgen05.ld/st — i.e. rows quad*32 .. +31 of the 128-row tile, thread == row, and
// column half (w >> 2): the two warps of a quadrant split the key columns of every tile between them, so two
// threads work on each row (twice the warps to hide instruction / memory latency).
// ---------------------------------------------------------------------------------------------
constexpr int ATT_THREADS = 256;

// Coalesced load of a [32 rows x 32 cols] fp32 bias block for one warp: each LDG.128 instruction covers 4 rows x 128
// contiguous bytes (4 wavefronts instead of 32 for a thread-per-row access), staged through a 4 KB XOR-swizzled smem
// tile, then every thread picks up the 32 values of ITS row.  `rows_left` clamps rows past the end of the tensor.
__device__ __forceinline__ void load_bias_block(const float* __restrict__ base, int64_t row_stride, int rows_left,
                                                uint32_t stage_addr, int lane, float (&out)[32]) {
  const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    const int rc = r < rows_left ? r : (rows_left - 1);
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + rc * row_stride) + ch);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr + r * 128 + ((ch ^ (r & 7)) << 4)), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(stage_addr + lane * 128 + ((c ^ (lane & 7)) << 4))
                 : "memory");
    out[4 * c] = v.x; out[4 * c + 1] = v.y; out[4 * c + 2] = v.z; out[4 * c + 3] = v.w;
  }
  __syncwarp();
}

// additive term (bias + mask) * log2e for the 32 columns [k0, k0+32) of query row q (k0 % 32 == 0)
__device__ __forceinline__ void additive32(const AttnParams& p, const float* bias_blk, int bias_rows_left,
                                           const float* mask_row, int k0, uint32_t stage_addr, int lane, float (&add)[32]) {
  if (bias_blk) {
    load_bias_block(bias_blk + k0, p.bias_q_stride, bias_rows_left, stage_addr, lane, add);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) add[j] = 0.f;
  }
  if (mask_row) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mask_row + k0 + j));
      add[j] += m.x; add[j + 1] += m.y; add[j + 2] += m.z; add[j + 3] += m.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) add[j] *= kLog2e;
}

int check_common(const X2kAttnArgs& a, const char* who) {
  X2K_REQUIRE(a.q && a.k && a.v && a.o && a.lse, "%s: NULL q/k/v/o/lse", who);
  X2K_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "%s: bad shape", who);
  X2K_REQUIRE(!a.kv_index || !a.kv_groups || a.n_kv > 0, "%s: kv_index / kv_groups need n_kv", who);
  X2K_REQUIRE(a.ld_q % 8 == 0 && a.ld_k % 8 == 0 && a.ld_v % 8 == 0 && a.ld_o % 8 == 0, "%s: ld must be multiples of 8", who);
  const int Lk_pad = (a.Lk + 15) & ~15;
  X2K_REQUIRE(!a.bias || (a.bias_q_stride % 4 == 0 && a.bias_h_stride % 4 == 0 && a.bias_q_stride >= ((a.Lk + 31) & ~31)),
              "%s: bias strides must be multiples of 4 and the row stride >= Lk rounded up to 32 (%d)", who, (a.Lk + 31) & ~31);
  X2K_REQUIRE(!a.mask || (a.mask_b_stride % 4 == 0 && a.mask_q_stride % 4 == 0 &&
                          (a.mask_q_stride == 0 ? a.mask_b_stride >= ((a.Lk + 31) & ~31) : a.mask_q_stride >= ((a.Lk + 31) & ~31))),
              "%s: mask strides must be multiples of 4 and cover Lk rounded up to 32 (%d)", who, (a.Lk + 31) & ~31);
  X2K_REQUIRE(a.dropout_p >= 0.f && a.dropout_p < 1.f, "%s: dropout_p", who);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  X2K_REQUIRE(al16(a.q) && al16(a.k) && al16(a.v) && al16(a.o) && al16(a.bias) && al16(a.mask), "%s: 16-byte alignment", who);
  return X2K_OK;
}

void fill_params(const X2kAttnArgs& a, AttnParams& p) {
  p.B = a.B; p.H = a.H; p.Lq = a.Lq; p.Lk = a.Lk; p.Lk_pad = (a.Lk + 15) & ~15;
  p.kv_index = a.kv_index;
  p.scale = a.scale; p.scale_log2 = a.scale * kLog2e;
  p.bias = a.bias; p.bias_h_stride = a.bias_h_stride; p.bias_q_stride = a.bias_q_stride;
  p.mask = a.mask; p.mask_b_stride = a.mask_b_stride; p.mask_q_stride = a.mask_q_stride;
  p.dropout_p = a.dropout_p; p.seed = a.dropout_seed; p.offset = a.dropout_offset; p.offset_dev = a.dropout_offset_dev;
  p.o = static_cast<__nv_bfloat16*>(a.o); p.ld_o = a.ld_o; p.lse = a.lse;
  p.d_o = static_cast<const __nv_bfloat16*>(a.d_o); p.ld_do = a.ld_do;
  p.dq = static_cast<__nv_bfloat16*>(a.dq); p.dk = static_cast<__nv_bfloat16*>(a.dk); p.dv = static_cast<__nv_bfloat16*>(a.dv);
  p.ld_dq = a.ld_dq; p.ld_dk = a.ld_dk; p.ld_dv = a.ld_dv;
  p.ds_out = static_cast<__nv_bfloat16*>(a.ds_out);
  p.ds_b_stride = a.ds_b_stride; p.ds_h_stride = a.ds_h_stride; p.ds_q_stride = a.ds_q_stride;
  p.delta = a.delta_ws;
}

}  // namespace
}  // namespace x2k
