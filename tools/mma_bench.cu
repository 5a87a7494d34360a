// mma_bench.cu — developer microbenchmark: cycles per tcgen05.mma (cta_group::1, M = 128, bf16) for the operand layouts the
// attention kernels use, one issuing thread, operands in shared memory.  Build: see tools/mma_bench.sh.
#include <cstdio>
#include <cuda_runtime.h>
#include "../x2vlm_b200/csrc/common.cuh"

using namespace x2k;

struct Case { int N, a_mn, b_mn, k_steps; const char* name; };

template <int MODE>  // 0: one thread inside a divergent branch; 1: whole warp (provably uniform index), MMA under elect.sync;
                    // 2: as 1, while the other three warps poll an mbarrier (mbar_wait_warp) the way waiting warps do
__global__ void __launch_bounds__(128, 1) bench_kernel(int N, int a_mn, int b_mn, int k_steps, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar_done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (MODE == 2 && warp_u != 0) mbar_wait_warp(&bar_done, 0);
  if (MODE == 0 ? threadIdx.x == 0 : warp_u == 0) {
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    // A: 128 x (16 * k_steps): K-major blocks [128][64] 16 KB each; MN-major: [k rows][64 of M], M chunks 16 KB apart
    const uint64_t fa = a_mn ? make_smem_desc(0, 16384, 1024) : make_smem_desc(0, 16, 1024);
    const uint64_t fb = b_mn ? make_smem_desc(0, 8192, 1024) : make_smem_desc(0, 16, 1024);
    const uint32_t a0 = sb, b0 = sb + 65536;
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < 16; ++i)
      for (int k = 0; k < k_steps; ++k) {
        const uint32_t aa = a_mn ? a0 + k * 2048 : a0 + (k >> 2) * 16384 + (k & 3) * 32;
        const uint32_t bb = b_mn ? b0 + k * 2048 : b0 + (k >> 2) * 32768 + (k & 3) * 32;
        if (MODE == 0 || elect_one()) umma_bf16(tmem, fa | ((aa & 0x3FFFF) >> 4), fb | ((bb & 0x3FFFF) >> 4), idesc, k != 0);
      }
      if (MODE == 0 || elect_one()) umma_commit(&bar);
      if (MODE >= 1) __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    if (MODE == 2 && threadIdx.x == 0) mbar_arrive(&bar_done);
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  const Case cases[] = {
      {128, 0, 0, 8, "S/dP   N=128 A K-major  B K-major "}, {256, 0, 0, 8, "       N=256 A K-major  B K-major "},
      {64, 0, 0, 8, "       N=64  A K-major  B K-major "},  {64, 0, 1, 8, "dQ     N=64  A K-major  B MN-major"},
      {64, 1, 1, 8, "dK/dV  N=64  A MN-major B MN-major"},  {128, 1, 1, 8, "       N=128 A MN-major B MN-major"},
      {64, 1, 0, 8, "       N=64  A MN-major B K-major "},  {32, 0, 0, 8, "       N=32  A K-major  B K-major "},
      {128, 0, 0, 4, "S (4)  N=128 A K-major  B K-major "},  {208, 0, 0, 4, "S (4)  N=208 A K-major  B K-major "}};
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(bench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 197632 + 1024);
  cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 197632 + 1024);
  cudaFuncSetAttribute(bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 197632 + 1024);
  for (int mode : {0, 1, 2}) {
    const int grid = 148;
    printf("mode %d (%s)\n", mode, mode == 0 ? "thread 0 in a divergent branch" : mode == 1 ? "warp-uniform issue code, elect.sync"
                                                                                              : "warp-uniform issue, 3 warps polling an mbarrier");
    for (const Case& c : cases) {
      const int reps = 50;
      if (mode == 0) bench_kernel<0><<<grid, 128, 197632 + 1024>>>(c.N, c.a_mn, c.b_mn, c.k_steps, reps, out);
      else if (mode == 1) bench_kernel<1><<<grid, 128, 197632 + 1024>>>(c.N, c.a_mn, c.b_mn, c.k_steps, reps, out);
      else bench_kernel<2><<<grid, 128, 197632 + 1024>>>(c.N, c.a_mn, c.b_mn, c.k_steps, reps, out);
      long long cyc = 0;
      cudaError_t e = cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      printf("  %s: %7.1f cycles per commit of %d MMAs = %6.1f / MMA (floor %d)\n", c.name, double(cyc) / reps, 16 * c.k_steps,
             double(cyc) / reps / c.k_steps / 16, 128 * c.N / 256);
    }
  }
  return 0;
}
