// runtime.cu — host-side plumbing of libx2k.so: error text, launch counter, device properties,
// TMA tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace x2k {

static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;  // immutable device property
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;  // immutable driver entry point
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return X2K_ERR_CUDA;
  }
  if (box_cols * 2 != 128 || box_rows > 256) {
    set_error("make_tmap: bad box %u x %u", box_rows, box_cols);
    return X2K_ERR_ARG;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (rows=%llu cols=%llu ld=%llu box=%ux%u base=%p)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, base);
    return X2K_ERR_CUDA;
  }
  return X2K_OK;
}

// fp32 row-major matrix [rows, cols] (ld elements), box = 32 rows x 16 columns (64 bytes), 64-byte swizzle: the shared-
// memory image of a box is exactly the GEMM epilogue's per-warp transpose tile (16-byte chunk c of row r at chunk
// c ^ ((r >> 1) & 3), gemm.cu epi_off) — residual tiles are TMA-loaded and fp32 output tiles TMA-stored in that layout.
int make_tmap_f32_epi(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return X2K_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 4};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 epilogue) failed: %d (rows=%llu cols=%llu ld=%llu base=%p)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, base);
    return X2K_ERR_CUDA;
  }
  return X2K_OK;
}

// 3-D view [n_seq][seq_rows][cols] of a row-major matrix whose sequences are seq_rows consecutive rows: box =
// [1][box_rows][64].  Rows of a box beyond seq_rows are out of bounds in dim 1: loads zero-fill them, stores skip
// them — a tile may be taller than the sequence without touching its neighbour.
int make_tmap_bf16_seq3d(CUtensorMap* tm, const void* base, uint64_t n_seq, uint64_t seq_rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return X2K_ERR_CUDA;
  }
  if (box_rows > 256 || box_rows == 0) {
    set_error("make_tmap_seq3d: bad box rows %u", box_rows);
    return X2K_ERR_ARG;
  }
  cuuint64_t gdim[3] = {cols, seq_rows, n_seq};
  cuuint64_t gstride[2] = {ld * 2, seq_rows * ld * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3d) failed: %d (n_seq=%llu rows=%llu cols=%llu ld=%llu box_rows=%u base=%p)", (int)r,
              (unsigned long long)n_seq, (unsigned long long)seq_rows, (unsigned long long)cols, (unsigned long long)ld,
              box_rows, base);
    return X2K_ERR_CUDA;
  }
  return X2K_OK;
}

int make_tmap_bf16_4d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t d2, uint64_t d3,
                      uint64_t row_stride, uint64_t d2_stride, uint64_t d3_stride) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return X2K_ERR_CUDA;
  }
  cuuint64_t gdim[4] = {cols, rows, d2, d3};
  cuuint64_t gstride[3] = {row_stride * 2, d2_stride * 2, d3_stride * 2};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4d) failed: %d (cols=%llu rows=%llu d2=%llu d3=%llu strides=%llu/%llu/%llu base=%p)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)d2, (unsigned long long)d3,
              (unsigned long long)row_stride, (unsigned long long)d2_stride, (unsigned long long)d3_stride, base);
    return X2K_ERR_CUDA;
  }
  return X2K_OK;
}

}  // namespace x2k

extern "C" int x2k_version(void) { return X2K_VERSION; }
extern "C" const char* x2k_last_error(void) { return x2k::g_err; }
extern "C" int64_t x2k_launch_count(void) { return x2k::g_launches.load(std::memory_order_relaxed); }
