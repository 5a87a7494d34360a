// mma_bench2.cu — developer microbenchmark: does switching the tcgen05.mma shape / operand layout / accumulator between
// back-to-back MMAs cost extra?  Replays the attention backward's per-tile sequence (dQ, dK, dV, S, dP) from one warp
// (uniform issue code, elect.sync) and compares it with the same number of MMAs of one kind.
#include <cstdio>
#include <cuda_runtime.h>
#include "../x2vlm_b200/csrc/common.cuh"

using namespace x2k;

struct Seg { int N, a_mn, b_mn, acc, count; };
struct Seq { Seg s[8]; int n; };

__global__ void __launch_bounds__(128, 1) bench_kernel(Seq q, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp_u == 0) {
    const uint32_t sb = smem_u32(smem);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int si = 0; si < q.n; ++si) {
        const Seg sg = q.s[si];
        const uint32_t idesc = make_idesc_bf16(128, sg.N, sg.a_mn, sg.b_mn);
        const uint64_t fa = sg.a_mn ? make_smem_desc(0, 16384, 1024) : make_smem_desc(0, 16, 1024);
        const uint64_t fb = sg.b_mn ? make_smem_desc(0, 8192, 1024) : make_smem_desc(0, 16, 1024);
        const uint32_t a0 = sb + (si & 1) * 32768, b0 = sb + 65536 + (si & 1) * 32768;
        for (int k = 0; k < sg.count; ++k) {
          const uint32_t aa = sg.a_mn ? a0 + k * 2048 : a0 + (k >> 2) * 16384 + (k & 3) * 32;
          const uint32_t bb = sg.b_mn ? b0 + k * 2048 : b0 + (k >> 2) * 32768 + (k & 3) * 32;
          if (elect_one()) umma_bf16(tmem + sg.acc, fa | ((aa & 0x3FFFF) >> 4), fb | ((bb & 0x3FFFF) >> 4), idesc, k != 0);
        }
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  struct Named { const char* name; Seq q; };
  const Named cases[] = {
      {"tile: dQ8 dK8 dV8 S4 dP4 (the kernel's order)        ", {{{64, 0, 1, 256, 8}, {64, 1, 1, 384, 8}, {64, 1, 1, 448, 8}, {128, 0, 0, 0, 4}, {128, 0, 0, 128, 4}}, 5}},
      {"same kinds, one accumulator                          ", {{{64, 0, 1, 256, 8}, {64, 1, 1, 256, 8}, {64, 1, 1, 256, 8}, {128, 0, 0, 0, 4}, {128, 0, 0, 0, 4}}, 5}},
      {"24 x N=64 MN/MN (3 chains, 3 accumulators) + 8 x N=128", {{{64, 1, 1, 256, 8}, {64, 1, 1, 384, 8}, {64, 1, 1, 448, 8}, {128, 0, 0, 0, 8}}, 4}},
      {"24 x N=64 MN/MN one chain + 8 x N=128 one chain       ", {{{64, 1, 1, 256, 24}, {128, 0, 0, 0, 8}}, 2}},
      {"32 x N=64 MN/MN one chain                             ", {{{64, 1, 1, 256, 32}}, 1}},
      {"8 chains of 4 x N=64 MN/MN, alternating accumulators  ", {{{64, 1, 1, 256, 4}, {64, 1, 1, 384, 4}, {64, 1, 1, 256, 4}, {64, 1, 1, 384, 4}, {64, 1, 1, 256, 4}, {64, 1, 1, 384, 4}, {64, 1, 1, 256, 4}, {64, 1, 1, 384, 4}}, 8}},
      {"alternating layouts KM/MN <-> MN/MN, chains of 4      ", {{{64, 0, 1, 256, 4}, {64, 1, 1, 384, 4}, {64, 0, 1, 256, 4}, {64, 1, 1, 384, 4}, {64, 0, 1, 256, 4}, {64, 1, 1, 384, 4}, {64, 0, 1, 256, 4}, {64, 1, 1, 384, 4}}, 8}},
  };
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 197632 + 1024);
  for (const Named& c : cases) {
    const int reps = 100;
    bench_kernel<<<148, 128, 197632 + 1024>>>(c.q, reps, out);
    long long cyc = 0;
    cudaError_t e = cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    int n = 0;
    for (int i = 0; i < c.q.n; ++i) n += c.q.s[i].count;
    printf("  %s: %7.1f cycles per commit of %d MMAs (incl. one commit + wait round trip)\n", c.name, double(cyc) / reps, n);
  }
  return 0;
}
