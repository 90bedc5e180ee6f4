// mma_bench.cu — microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32) by shape / operand major / A source.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I tae_b200/csrc tools/mma_bench.cu -o gpurun_out/mma_bench -lcuda
// One thread issues `count` MMAs back to back, commits, waits: (t1 - t0) / count = cycles per MMA at saturation.
#include <cstdio>
#include <cstdlib>

#include "sm100.cuh"

namespace tae {
std::atomic<uint64_t> g_launch_count{0};
void set_error(const char*, ...) {}
int num_sms() { return 148; }
}  // namespace tae

using namespace tae;
using namespace tae::sm100;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

struct Cfg {
  int m, n, a_mn, b_mn, a_tmem, count;
};

__global__ void __launch_bounds__(128, 1) bench(const Cfg* cfgs, int ncfg, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      uint32_t phase = 0;
      const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 65536);
      for (int c = 0; c < ncfg; ++c) {
        const Cfg cf = cfgs[c];
        const uint32_t idesc = make_idesc_bf16(cf.m, cf.n, cf.a_mn, cf.b_mn);
        const uint32_t a_lo = cf.a_mn ? desc_lo(sA, 8192) : desc_lo(sA), b_lo = cf.b_mn ? desc_lo(sB, 8192) : desc_lo(sB);
        const uint32_t a_st = cf.a_mn ? (2048 >> 4) : 2, b_st = cf.b_mn ? (2048 >> 4) : 2;
        for (int rep = 0; rep < 3; ++rep) {
          const long long t0 = clock64();
          for (int i = 0; i < cf.count; i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (cf.a_tmem)
                umma_f16_ts(tmem, tmem + 256 + (k & 3) * 8, make_smem_desc(sB + (k & 3) * 32, 0, 1024), idesc, 1);
              else
                umma_f16_lo(tmem, a_lo + (k & 3) * a_st, b_lo + (k & 3) * b_st, idesc, 1);
            }
          }
          umma_commit(&bar);
          mbar_wait(&bar, phase);
          phase ^= 1;
          const long long t1 = clock64();
          if (blockIdx.x == 0) out[c * 3 + rep] = t1 - t0;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  Cfg h[] = {
      {128, 256, 0, 0, 0, 64}, {128, 128, 0, 0, 0, 64}, {128, 64, 0, 0, 0, 64}, {128, 32, 0, 0, 0, 64},
      {128, 64, 0, 1, 0, 64},  {128, 64, 1, 1, 0, 64},  {64, 64, 1, 1, 0, 64},  {64, 64, 0, 0, 0, 64},
      {128, 128, 0, 1, 0, 64}, {128, 128, 1, 1, 0, 64}, {128, 256, 1, 1, 0, 64}, {64, 128, 0, 0, 0, 64},
      {64, 256, 0, 0, 0, 64},  {128, 64, 0, 0, 1, 64},  {128, 64, 0, 1, 1, 64},  {128, 128, 0, 0, 1, 64},
      {128, 256, 0, 0, 1, 64}, {128, 64, 0, 0, 0, 8},   {128, 64, 0, 0, 0, 16},   {128, 256, 0, 0, 0, 8},
  };
  const int n = sizeof(h) / sizeof(h[0]);
  Cfg* d;
  long long* out;
  cudaMalloc(&d, sizeof(h));
  cudaMalloc(&out, n * 3 * sizeof(long long));
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 148}) {
    cudaMemset(out, 0, n * 3 * sizeof(long long));
    bench<<<grid, 128, 200 * 1024>>>(d, n, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("grid %d: CUDA error %s\n", grid, cudaGetErrorString(e));
      return 1;
    }
    long long r[64 * 3];
    cudaMemcpy(r, out, n * 3 * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("grid=%d\n", grid);
    for (int c = 0; c < n; ++c)
      printf("  M=%3d N=%3d a_mn=%d b_mn=%d a_tmem=%d count=%2d: total %6lld cyc  -> %.1f cyc/MMA (reps %lld %lld %lld)\n", h[c].m,
             h[c].n, h[c].a_mn, h[c].b_mn, h[c].a_tmem, h[c].count, r[c * 3 + 2], (double)r[c * 3 + 2] / h[c].count, r[c * 3],
             r[c * 3 + 1], r[c * 3 + 2]);
  }
  return 0;
}
