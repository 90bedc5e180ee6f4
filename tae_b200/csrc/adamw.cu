// adamw.cu — fused AdamW over a flat fp32 arena (HBM-bound: 28 B/param + 2 B/param bf16 shadow).
//
// Replaces torch.optim.AdamW(param_groups, lr, betas=(0.9, 0.95), fused=True) (train.py:109) and the
// GradScaler / grad-norm passes around it (util/misc.py:252-262, :274-286): one 128-bit vectorised pass reads
// p, g, m, v, writes p, m, v and the bf16 shadow copy the GEMMs consume, optionally accumulating sum(g^2)
// and honouring a device-side found_inf flag (GradScaler.step skip semantics).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace tae {
namespace opt {

struct AdamArgs {
  float lr_wd_mul;    // 1 - lr * wd
  float beta1, beta2;
  float one_m_beta1, one_m_beta2;
  float step_size;    // lr / (1 - beta1^t)
  float inv_bc2_sqrt;  // 1 / sqrt(1 - beta2^t)
  float eps;
  float grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  p *= a.lr_wd_mul;
  m = m + (g - m) * a.one_m_beta1;  // lerp, as torch's fused kernel
  v = a.beta2 * v + a.one_m_beta2 * g * g;
  const float denom = sqrtf(v) * a.inv_bc2_sqrt + a.eps;
  p -= a.step_size * (m / denom);
}

// `dev_args` != nullptr: the step's scalars live in device memory (9 floats written by tae_adamw_hyper + a copy), so a
// captured CUDA graph of the training step can be replayed with a new lr / step count without re-capturing.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ pb, size_t n, AdamArgs a, const AdamArgs* __restrict__ dev_args, float* grad_sq_sum,
             const int* __restrict__ found_inf) {
  if (found_inf != nullptr && *found_inf != 0) return;
  if (dev_args != nullptr) a = *dev_args;
  const size_t n4 = n / 4;
  float sq = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = ld_nc_f4(g + i * 4);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    gv.x *= a.grad_scale; gv.y *= a.grad_scale; gv.z *= a.grad_scale; gv.w *= a.grad_scale;
    sq += (gv.x * gv.x + gv.y * gv.y) + (gv.z * gv.z + gv.w * gv.w);
    adam_one(pv.x, gv.x, mv.x, vv.x, a);
    adam_one(pv.y, gv.y, mv.y, vv.y, a);
    adam_one(pv.z, gv.z, mv.z, vv.z, a);
    adam_one(pv.w, gv.w, mv.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(pv.x, pv.y);
      o.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(pb)[i] = o;
    }
  }
  // tail (n % 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = n4 * 4 + threadIdx.x;
    float pv = p[i], gv = g[i] * a.grad_scale, mv = m[i], vv = v[i];
    sq += gv * gv;
    adam_one(pv, gv, mv, vv, a);
    p[i] = pv;
    m[i] = mv;
    v[i] = vv;
    if (pb != nullptr) pb[i] = __float2bfloat16_rn(pv);
  }
  if (grad_sq_sum != nullptr) {
    __shared__ float red[8];
    sq = warp_sum(sq);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    if (warp == 0) {
      float s = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
      s = warp_sum(s);
      if (lane == 0) atomicAdd(grad_sq_sum, s);
    }
  }
}

__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = ld_nc_f4(src + i * 8), b = ld_nc_f4(src + i * 8 + 4);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y);
    o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t i = n8 * 8 + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

__global__ void __launch_bounds__(256)
grad_stats_kernel(const float* __restrict__ g, size_t n, float* sq_sum, int* found_inf) {
  float sq = 0.f;
  int bad = 0;
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = ld_nc_f4(g + i * 4);
    sq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[n4 * 4 + threadIdx.x];
    sq += v * v;
    bad |= !isfinite(v);
  }
  __shared__ float red[8];
  __shared__ int sbad;
  if (threadIdx.x == 0) sbad = 0;
  __syncthreads();
  if (bad) atomicOr(&sbad, 1);
  sq = warp_sum(sq);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = sq;
  __syncthreads();
  if (warp == 0) {
    float s = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) {
      if (sq_sum != nullptr) atomicAdd(sq_sum, s);
      if (found_inf != nullptr && sbad) atomicOr(found_inf, 1);
    }
  }
}

static int flat_grid(size_t nvec) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  size_t blocks = (nvec + 255) / 256;
  const size_t cap = (size_t)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace opt
}  // namespace tae

static_assert(sizeof(tae::opt::AdamArgs) == TAE_ADAMW_HYPER_FLOATS * sizeof(float), "tae_adamw_hyper layout");

extern "C" int tae_adamw_hyper(float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                               float grad_scale, float* out) {
  using namespace tae;
  using namespace tae::opt;
  TAE_CHECK_SHAPE(out != nullptr, "tae_adamw_hyper: NULL output");
  TAE_CHECK_SHAPE(step >= 1, "tae_adamw_hyper: step must be >= 1 (got %d)", step);
  AdamArgs a;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.lr_wd_mul = (float)(1.0 - (double)lr * (double)weight_decay);
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.one_m_beta1 = 1.0f - beta1;
  a.one_m_beta2 = 1.0f - beta2;
  a.step_size = (float)((double)lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.eps = eps;
  a.grad_scale = grad_scale;
  memcpy(out, &a, sizeof(a));
  return TAE_OK;
}

static int adamw_launch(float* p, const float* g, float* m, float* v, tae_bf16* p_bf16, size_t n,
                        const tae::opt::AdamArgs& a, const float* dev_hyper, float* grad_sq_sum,
                        const int32_t* found_inf, void* stream_) {
  using namespace tae;
  using namespace tae::opt;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(p && g && m && v, "tae_adamw_step: NULL arena");
  TAE_CHECK_SHAPE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(dev_hyper) & 3) == 0,
                  "tae_adamw_step: arenas must be 16-byte aligned");
  adamw_kernel<<<flat_grid(n / 4 + 1), 256, 0, stream>>>(p, g, m, v, reinterpret_cast<bf16*>(p_bf16), n, a,
                                                          reinterpret_cast<const AdamArgs*>(dev_hyper), grad_sq_sum,
                                                          reinterpret_cast<const int*>(found_inf));
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_adamw_step(float* p, const float* g, float* m, float* v, tae_bf16* p_bf16, size_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                              float* grad_sq_sum, const int32_t* found_inf, void* stream_) {
  if (n == 0) return TAE_OK;
  tae::opt::AdamArgs a;
  const int rc = tae_adamw_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale, reinterpret_cast<float*>(&a));
  if (rc != TAE_OK) return rc;
  return adamw_launch(p, g, m, v, p_bf16, n, a, nullptr, grad_sq_sum, found_inf, stream_);
}

extern "C" int tae_adamw_step_dev(float* p, const float* g, float* m, float* v, tae_bf16* p_bf16, size_t n,
                                  const float* hyper_dev, float* grad_sq_sum, const int32_t* found_inf, void* stream_) {
  using namespace tae;
  if (n == 0) return TAE_OK;
  TAE_CHECK_SHAPE(hyper_dev != nullptr, "tae_adamw_step_dev: NULL hyper-parameter block");
  tae::opt::AdamArgs a = {};
  return adamw_launch(p, g, m, v, p_bf16, n, a, hyper_dev, grad_sq_sum, found_inf, stream_);
}

extern "C" int tae_cast_f32_to_bf16(const float* src, tae_bf16* dst, size_t n, void* stream_) {
  using namespace tae;
  using namespace tae::opt;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n == 0) return TAE_OK;
  TAE_CHECK_SHAPE(src && dst, "tae_cast_f32_to_bf16: NULL pointer");
  TAE_CHECK_SHAPE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                  "tae_cast_f32_to_bf16: pointers must be 16-byte aligned");
  cast_kernel<<<flat_grid(n / 8 + 1), 256, 0, stream>>>(src, reinterpret_cast<bf16*>(dst), n);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_grad_stats(const float* g, size_t n, float* sq_sum, int32_t* found_inf, void* stream_) {
  using namespace tae;
  using namespace tae::opt;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n == 0) return TAE_OK;
  TAE_CHECK_SHAPE(g != nullptr && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "tae_grad_stats: g must be 16-byte aligned");
  grad_stats_kernel<<<flat_grid(n / 4 + 1), 256, 0, stream>>>(g, n, sq_sum, reinterpret_cast<int*>(found_inf));
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
