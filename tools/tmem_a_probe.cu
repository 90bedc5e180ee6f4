// tmem_a_probe.cu — determines the TMEM layout of a bf16 A operand for tcgen05.mma (A-from-TMEM form).
// A cells are written with tcgen05.st (lane = row m, 32-bit column c = packed pair), B = 16x16 identity in smem
// (K-major, 128B swizzle), D[m, n] = A(m, k = n).  Prints D for a few rows.
#include <cstdio>
#include "sm100.cuh"
namespace tae { std::atomic<uint64_t> g_launch_count{0}; void set_error(const char*, ...) {} int num_sms() { return 148; } }
using namespace tae;
using namespace tae::sm100;

__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) probe(float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B: [16 rows (n) x 64 cols (k)] bf16, K-major, 128B swizzle: B[n][k] = (n == k)
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int n = i / 64, kk = i % 64;
    const uint32_t off = n * 128 + (((kk / 8) ^ (n & 7)) << 4) + (kk % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = __float2bfloat16(n == kk ? 1.0f : 0.0f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  // A cells at columns 256..263: lane r = warp*32 + lane; cell c holds (lo, hi) = (2c + 64*(r&1), 2c + 1 + 64*(r&1))
  {
    const int r = warp * 32 + lane;
    uint32_t v[8];
    for (int c = 0; c < 8; ++c) {
      const float lo = 2 * c + 64 * (r & 1) + 0.5f * (r >> 6), hi = lo + 1;
      __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
      v[c] = *reinterpret_cast<uint32_t*>(&t);
    }
    const uint32_t taddr = tmem + 256 + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
    umma_f16_ts(tmem, tmem + 256, make_smem_desc(smem_u32(smem), 0, 1024), idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  {
    uint32_t d[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                   "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]) : "r"(taddr));
    tmem_ld_wait();
    const int r = warp * 32 + lane;
    for (int n = 0; n < 16; ++n) out[r * 16 + n] = __uint_as_float(d[n]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  probe<<<1, 128, 8192>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  float h[128 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int r : {0, 1, 2, 33, 64, 127}) {
    printf("row %3d:", r);
    for (int n = 0; n < 16; ++n) printf(" %5.1f", h[r * 16 + n]);
    printf("\n");
  }
  printf("expected if cell c = (A[m][2c], A[m][2c+1]): row r -> n + 64*(r&1) + 0.5*(r>>6)\n");
  return 0;
}
