// fp32_mode.cu — the fp32 ("no autocast") numerics mode of the hot path: same call sites as the bf16 kernels
// (tae.py:46-54, 72-82, 100-105, 128-131, 256-265), every activation kept in fp32.
//
// GEMMs still run on the tcgen05 tensor cores: tae_split3_bf16 splits an fp32 operand into three bf16 terms
// (x = hi + mid + lo exactly: 3 x 8 mantissa bits), and the host issues the six significant cross products
// (lo*hi, hi*lo, mid*mid, mid*hi, hi*mid, hi*hi) as tae_gemm launches accumulating into one fp32 output
// (TAE_EPI_F32_ACC, beta = 1).  Products of bf16 pairs are exact in fp32, so the result carries fp32-level error
// (dropped terms are <= 2^-24 relative).  The kernels in this file are the element-wise / row-wise / attention
// pieces around those GEMMs.  They are written for clarity, not speed: this mode exists for the 1e-4 parity gate.
#include "common.cuh"

namespace tae {
namespace f32 {

__global__ void split3_kernel(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ mid,
                              bf16* __restrict__ lo, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const bf16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);  // exact
    const bf16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);  // exact
    hi[i] = h;
    mid[i] = m;
    lo[i] = __float2bfloat16_rn(r2);
  }
}

// y[m, n] = acc[m, n] + bias[n] + resid[m % resid_rows, n];  optionally act[m, n] = gelu_erf(y[m, n])
__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, const float* __restrict__ resid,
                                int resid_rows, float* __restrict__ act, size_t M, int N) {
  const size_t total = M * (size_t)N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / N;
    const int n = (int)(i - m * N);
    float v = y[i];
    if (bias != nullptr) v += bias[n];
    if (resid != nullptr) v += resid[(m % (size_t)resid_rows) * N + n];
    y[i] = v;
    if (act != nullptr) act[i] = gelu_erf(v);
  }
}

__global__ void gelu_bwd_kernel(const float* __restrict__ h, const float* __restrict__ da, float* __restrict__ dh,
                                size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dh[i] = da[i] * gelu_erf_grad(h[i]);
}

// out = a + b (b may be NULL -> copy)
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = a[i] + (b != nullptr ? b[i] : 0.f);
}

// ---- LayerNorm, one warp per row --------------------------------------------------------------
__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int rows, int D,
                              float eps) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const float* xr = x + (size_t)row * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mu = warp_sum(s) / (float)D;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = xr[c] - mu;
      q += d * d;
    }
    const float rs = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
    float* yr = y + (size_t)row * D;
    for (int c = lane; c < D; c += 32) yr[c] = (xr[c] - mu) * rs * gamma[c] + beta[c];
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dres_out = dx + dres_in
__global__ void ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* dres_in,
                              float* dres_out, int rows, int D) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const size_t base = (size_t)row * D;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float g = dy[base + c] * gamma[c];
      s1 += g;
      s2 += g * ((x[base + c] - mu) * rs);
    }
    const float c1 = warp_sum(s1) / (float)D, c2 = warp_sum(s2) / (float)D;
    for (int c = lane; c < D; c += 32) {
      const float g = dy[base + c] * gamma[c];
      const float xh = (x[base + c] - mu) * rs;
      const float dx = rs * (g - c1 - xh * c2);
      dres_out[base + c] = dx + (dres_in != nullptr ? dres_in[base + c] : 0.f);
    }
  }
}

// dgamma[c] (+)= sum_r dy[r,c] * xhat[r,c];  dbeta[c] (+)= sum_r dy[r,c].  One thread per column (coalesced rows).
__global__ void ln_param_grads_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                      const float* __restrict__ mean, const float* __restrict__ rstd, float* dgamma,
                                      float* dbeta, int rows, int D, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float sg = 0.f, sb = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float d = dy[(size_t)r * D + c];
    sg += d * ((x[(size_t)r * D + c] - mean[r]) * rstd[r]);
    sb += d;
  }
  if (dgamma != nullptr) dgamma[c] = (accumulate & 1) ? dgamma[c] + sg : sg;
  if (dbeta != nullptr) dbeta[c] = (accumulate & 2) ? dbeta[c] + sb : sb;
}

// ---- attention, one warp per query / key row ---------------------------------------------------
// qkv fp32 [B*N, 3*H*hd], out fp32 [B*N, H*hd], lse [B, H, N]; prob: per-warp scratch row in shared memory.
__global__ void attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse, int B,
                                int N, int H, int hd, float scale) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* prob = sm + (size_t)w * N;
  const int D = H * hd;
  const size_t ld = (size_t)3 * D;
  const int total = B * H * N;
  for (int item = blockIdx.x * wpb + w; item < total; item += gridDim.x * wpb) {
    const int i = item % N, bh = item / N, h = bh % H, b = bh / H;
    const float* q = qkv + ((size_t)b * N + i) * ld + (size_t)h * hd;
    const float* kbase = qkv + (size_t)b * N * ld + D + (size_t)h * hd;
    const float* vbase = kbase + D;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      const float* k = kbase + (size_t)j * ld;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(q[d], k[d], s);
      s *= scale;
      prob[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float l = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float p = expf(prob[j] - mx);
      prob[j] = p;
      l += p;
    }
    l = warp_sum(l);
    __syncwarp();
    if (lane == 0) lse[item] = mx + logf(l);
    const float inv = 1.0f / l;
    float* o = out + ((size_t)b * N + i) * D + (size_t)h * hd;
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(prob[j], vbase[(size_t)j * ld + d], acc);
      o[d] = acc * inv;
    }
    __syncwarp();
  }
}

// pass 1 (per query row): P, dS rows to the workspace; dQ_i = sum_j dS_ij K_j
__global__ void attn_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ out,
                                  const float* __restrict__ dout, const float* __restrict__ lse, float* __restrict__ dqkv,
                                  float* __restrict__ wsP, float* __restrict__ wsdS, int B, int N, int H, int hd,
                                  float scale) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int D = H * hd;
  const size_t ld = (size_t)3 * D;
  const int total = B * H * N;
  for (int item = blockIdx.x * wpb + w; item < total; item += gridDim.x * wpb) {
    const int i = item % N, bh = item / N, h = bh % H, b = bh / H;
    const float* q = qkv + ((size_t)b * N + i) * ld + (size_t)h * hd;
    const float* kbase = qkv + (size_t)b * N * ld + D + (size_t)h * hd;
    const float* vbase = kbase + D;
    const float* o = out + ((size_t)b * N + i) * D + (size_t)h * hd;
    const float* go = dout + ((size_t)b * N + i) * D + (size_t)h * hd;
    float dl = 0.f;
    for (int d = lane; d < hd; d += 32) dl += o[d] * go[d];
    const float delta = warp_sum(dl);
    const float L = lse[item];
    float* prow = wsP + (size_t)item * N;
    float* srow = wsdS + (size_t)item * N;
    for (int j = lane; j < N; j += 32) {
      const float* k = kbase + (size_t)j * ld;
      const float* v = vbase + (size_t)j * ld;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) {
        s = fmaf(q[d], k[d], s);
        dp = fmaf(go[d], v[d], dp);
      }
      const float p = expf(s * scale - L);
      prow[j] = p;
      srow[j] = p * (dp - delta) * scale;
    }
    __syncwarp();
    float* dq = dqkv + ((size_t)b * N + i) * ld + (size_t)h * hd;
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc = fmaf(srow[j], kbase[(size_t)j * ld + d], acc);
      dq[d] = acc;
    }
    __syncwarp();
  }
}

// pass 2 (per key row): dV_j = sum_i P_ij dO_i;  dK_j = sum_i dS_ij Q_i
__global__ void attn_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, float* __restrict__ dqkv,
                                   const float* __restrict__ wsP, const float* __restrict__ wsdS, int B, int N, int H,
                                   int hd) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int D = H * hd;
  const size_t ld = (size_t)3 * D;
  const int total = B * H * N;
  for (int item = blockIdx.x * wpb + w; item < total; item += gridDim.x * wpb) {
    const int j = item % N, bh = item / N, h = bh % H, b = bh / H;
    const float* qbase = qkv + (size_t)b * N * ld + (size_t)h * hd;
    const float* gobase = dout + (size_t)b * N * D + (size_t)h * hd;
    const float* pcol = wsP + (size_t)bh * N * N + j;
    const float* scol = wsdS + (size_t)bh * N * N + j;
    float* dk = dqkv + ((size_t)b * N + j) * ld + D + (size_t)h * hd;
    float* dv = dk + D;
    for (int d = lane; d < hd; d += 32) {
      float av = 0.f, ak = 0.f;
      for (int i = 0; i < N; ++i) {
        av = fmaf(pcol[(size_t)i * N], gobase[(size_t)i * D + d], av);
        ak = fmaf(scol[(size_t)i * N], qbase[(size_t)i * ld + d], ak);
      }
      dv[d] = av;
      dk[d] = ak;
    }
  }
}

// ---- im2col / loss in fp32 ----------------------------------------------------------------------
__global__ void im2col_kernel(const float* __restrict__ imgs, float* __restrict__ cols, int B, int S, int p) {
  const int g = S / p, Kp = 3 * p * p;
  const size_t total = (size_t)B * 3 * S * S;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % S);
    size_t r = idx / S;
    const int y = (int)(r % S);
    r /= S;
    const int c = (int)(r % 3), b = (int)(r / 3);
    const int h = y / p, i = y - h * p, w = x / p, j = x - w * p;
    cols[((size_t)b * g * g + (size_t)h * g + w) * Kp + (size_t)c * p * p + i * p + j] = imgs[idx];
  }
}

// loss_accum += sum (pred - patchify(imgs))^2 / numel;  dpred = grad_scale * 2 (pred - target) / numel
__global__ void mse_loss_kernel(const float* __restrict__ pred, const float* __restrict__ imgs, float* loss_accum,
                                float* __restrict__ dpred, const float* __restrict__ grad_scale, int B, int S, int p) {
  const int g = S / p, Kp = 3 * p * p;
  const size_t per_img = (size_t)3 * S * S, total = (size_t)B * per_img;
  const float inv = 1.0f / (float)total;
  const float gs = (grad_scale != nullptr ? grad_scale[0] : 1.0f) * 2.0f * inv;
  float lsum = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per_img);
    const size_t r = idx - (size_t)b * per_img;
    const int n = (int)(r / Kp), e = (int)(r - (size_t)n * Kp);
    const int c = e % 3, ij = e / 3, i = ij / p, j = ij - i * p, h = n / g, w = n - h * g;
    const float t = imgs[(((size_t)b * 3 + c) * S + (size_t)h * p + i) * S + (size_t)w * p + j];
    const float d = pred[idx] - t;
    lsum = fmaf(d, d, lsum);
    if (dpred != nullptr) dpred[idx] = d * gs;
  }
  lsum = warp_sum(lsum);
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = lsum;
  __syncthreads();
  if (w == 0) {
    float s = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(loss_accum, s * inv);
  }
}

static int grid_1d(size_t n, int threads) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  size_t need = (n + threads - 1) / threads;
  const size_t cap = (size_t)sms * 8;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

}  // namespace f32
}  // namespace tae

using namespace tae;
using namespace tae::f32;

extern "C" int tae_split3_bf16(const float* x, tae_bf16* hi, tae_bf16* mid, tae_bf16* lo, size_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(x && hi && mid && lo && n > 0, "tae_split3_bf16: bad arguments");
  split3_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(x, reinterpret_cast<bf16*>(hi), reinterpret_cast<bf16*>(mid),
                                                      reinterpret_cast<bf16*>(lo), n);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_bias_act_f32(float* y, const float* bias, const float* resid, int32_t resid_rows, float* act,
                                int32_t M, int32_t N, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(y && M > 0 && N > 0, "tae_bias_act_f32: bad arguments");
  TAE_CHECK_SHAPE(resid == nullptr || resid_rows > 0, "tae_bias_act_f32: resid needs resid_rows > 0");
  bias_act_kernel<<<grid_1d((size_t)M * N, 256), 256, 0, stream>>>(y, bias, resid, resid_rows > 0 ? resid_rows : 1, act,
                                                                    (size_t)M, N);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_gelu_bwd_f32(const float* h, const float* da, float* dh, size_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(h && da && dh && n > 0, "tae_gelu_bwd_f32: bad arguments");
  gelu_bwd_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(h, da, dh, n);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_add_f32(const float* a, const float* b, float* out, size_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(a && out && n > 0, "tae_add_f32: bad arguments");
  add_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(a, b, out, n);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                     float* rstd, int32_t rows, int32_t D, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(rows > 0 && D > 0, "tae_layernorm_fwd_f32: rows/D must be positive");
  ln_fwd_kernel<<<grid_1d((size_t)rows * 32, 256), 256, 0, stream>>>(x, gamma, beta, y, mean, rstd, rows, D, eps);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_layernorm_bwd_f32(const float* dy, const float* x, const float* mean, const float* rstd,
                                     const float* gamma, const float* dres_in, float* dres_out, float* dgamma,
                                     float* dbeta, int32_t accumulate, int32_t rows, int32_t D, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(rows > 0 && D > 0 && dres_out != nullptr, "tae_layernorm_bwd_f32: bad arguments");
  ln_bwd_kernel<<<grid_1d((size_t)rows * 32, 256), 256, 0, stream>>>(dy, x, mean, rstd, gamma, dres_in, dres_out, rows, D);
  TAE_CHECK_LAUNCH();
  if (dgamma != nullptr || dbeta != nullptr) {
    ln_param_grads_kernel<<<(D + 127) / 128, 128, 0, stream>>>(dy, x, mean, rstd, dgamma, dbeta, rows, D, accumulate);
    TAE_CHECK_LAUNCH();
  }
  return TAE_OK;
}

extern "C" int tae_attention_fwd_f32(const float* qkv, float* out, float* lse, int32_t B, int32_t N, int32_t H,
                                     int32_t hd, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && N > 0 && N <= 4096 && H > 0 && hd > 0, "tae_attention_fwd_f32: bad shape");
  const int wpb = 8;
  const size_t smem = (size_t)wpb * N * sizeof(float);
  TAE_CHECK_SHAPE(smem <= 48 * 1024, "tae_attention_fwd_f32: N=%d too long", N);
  attn_fwd_kernel<<<grid_1d((size_t)B * H * N * 32, 256), 256, smem, stream>>>(qkv, out, lse, B, N, H, hd,
                                                                                1.0f / sqrtf((float)hd));
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" size_t tae_attention_bwd_f32_workspace_floats(int32_t B, int32_t N, int32_t H) {
  return (size_t)2 * B * H * N * N;
}

extern "C" int tae_attention_bwd_f32(const float* qkv, const float* out, const float* dout, const float* lse,
                                     float* dqkv, float* workspace, int32_t B, int32_t N, int32_t H, int32_t hd,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && N > 0 && H > 0 && hd > 0 && workspace != nullptr, "tae_attention_bwd_f32: bad arguments");
  float* wsP = workspace;
  float* wsdS = workspace + (size_t)B * H * N * N;
  const int grid = grid_1d((size_t)B * H * N * 32, 256);
  const float scale = 1.0f / sqrtf((float)hd);
  attn_bwd_q_kernel<<<grid, 256, 0, stream>>>(qkv, out, dout, lse, dqkv, wsP, wsdS, B, N, H, hd, scale);
  TAE_CHECK_LAUNCH();
  attn_bwd_kv_kernel<<<grid, 256, 0, stream>>>(qkv, dout, dqkv, wsP, wsdS, B, N, H, hd);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_im2col_f32(const float* imgs, float* cols, int32_t B, int32_t S, int32_t p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && S > 0 && p > 0 && S % p == 0, "tae_im2col_f32: bad shape B=%d S=%d p=%d", B, S, p);
  tae::f32::im2col_kernel<<<grid_1d((size_t)B * 3 * S * S, 256), 256, 0, stream>>>(imgs, cols, B, S, p);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_mse_loss_f32(const float* pred, const float* imgs, float* loss_accum, float* dpred,
                                const float* grad_scale, int32_t B, int32_t S, int32_t p, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && S > 0 && p > 0 && S % p == 0 && loss_accum != nullptr, "tae_mse_loss_f32: bad arguments");
  tae::f32::mse_loss_kernel<<<grid_1d((size_t)B * 3 * S * S, 256), 256, 0, stream>>>(pred, imgs, loss_accum, dpred,
                                                                                     grad_scale, B, S, p);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
