// layernorm.cu — LayerNorm forward / backward for the fp32 residual stream (HBM-bound kernels).
//
// Reference call sites: nn.LayerNorm(D, eps=1e-6) as norm1/norm2/norm/decoder_norm
// (tae.py:122,125,159,168; eps via partial at tae.py:435).  Under autocast LayerNorm runs in fp32 and the
// following nn.Linear re-casts its output to the low-precision dtype; here the normalised row is written
// directly as bf16 (same rounding point).  Statistics: fp32, biased variance.
//
// Thread mapping: a thread owns VPT fixed groups of 4 consecutive columns for ALL rows its CTA visits
// (gamma/beta live in registers; the backward column sums dgamma/dbeta/colsum accumulate in registers with
// no atomics); row statistics are CTA-wide reductions, batched over ROWS rows per iteration so that ROWS
// independent 128-bit loads per thread are in flight.
#include "common.cuh"

namespace tae {
namespace ln {

constexpr int ROWS = 4;       // rows per CTA iteration
constexpr int MAX_WARPS = 16;  // blockDim <= 512

// CTA-wide sum of NV values per thread (all threads get the result). red: smem [NV][MAX_WARPS].
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // protect red from the previous use
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * MAX_WARPS + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = (lane < nwarps) ? red[i * MAX_WARPS + lane] : 0.0f;
    v[i] = warp_sum(s);
  }
}

template <int VPT>
__global__ void __launch_bounds__(512)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D,
              float eps) {
  __shared__ float red[2 * ROWS * MAX_WARPS];
  const int tid = threadIdx.x, nthr = blockDim.x;
  float4 g[VPT], b[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int col = (v * nthr + tid) * 4;
    g[v] = *reinterpret_cast<const float4*>(gamma + col);
    b[v] = *reinterpret_cast<const float4*>(beta + col);
  }
  const float invD = 1.0f / (float)D;
  for (int r0 = blockIdx.x * ROWS; r0 < rows; r0 += gridDim.x * ROWS) {
    float4 xv[ROWS][VPT];
    float s[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      s[r] = 0.0f;
      const int row = r0 + r;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        if (row < rows) {
          xv[r][v] = ld_nc_f4(x + (size_t)row * D + (v * nthr + tid) * 4);
        } else {
          xv[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        s[r] += (xv[r][v].x + xv[r][v].y) + (xv[r][v].z + xv[r][v].w);
      }
    }
    block_sum<ROWS>(s, red);
    float mu[ROWS], q[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      mu[r] = s[r] * invD;
      q[r] = 0.0f;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const float dx = xv[r][v].x - mu[r], dy = xv[r][v].y - mu[r], dz = xv[r][v].z - mu[r], dw = xv[r][v].w - mu[r];
        q[r] += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    block_sum<ROWS>(q, red);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = r0 + r;
      if (row >= rows) continue;
      const float rs = rsqrtf(q[r] * invD + eps);
      if (tid == 0) {
        mean_out[row] = mu[r];
        rstd_out[row] = rs;
      }
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const float o0 = (xv[r][v].x - mu[r]) * rs * g[v].x + b[v].x;
        const float o1 = (xv[r][v].y - mu[r]) * rs * g[v].y + b[v].y;
        const float o2 = (xv[r][v].z - mu[r]) * rs * g[v].z + b[v].z;
        const float o3 = (xv[r][v].w - mu[r]) * rs * g[v].w + b[v].w;
        uint2 o;
        o.x = pack_bf16x2(o0, o1);
        o.y = pack_bf16x2(o2, o3);
        *reinterpret_cast<uint2*>(y + (size_t)row * D + (v * nthr + tid) * 4) = o;
      }
    }
  }
}

// Backward.  partials layout: [gridDim.x][3][D]  (0: dgamma, 1: dbeta, 2: colsum(bf16(dres_out)))
template <int VPT>
__global__ void __launch_bounds__(512)
ln_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* dres_in,
              float* dres_out, bf16* __restrict__ dres_out_b, float* __restrict__ partials, int rows, int D) {
  __shared__ float red[2 * ROWS * MAX_WARPS];
  const int tid = threadIdx.x, nthr = blockDim.x;
  float4 g[VPT], acc_dg[VPT], acc_db[VPT], acc_cs[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    g[v] = *reinterpret_cast<const float4*>(gamma + (v * nthr + tid) * 4);
    acc_dg[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_db[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_cs[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invD = 1.0f / (float)D;
  for (int r0 = blockIdx.x * ROWS; r0 < rows; r0 += gridDim.x * ROWS) {
    float4 xh[ROWS][VPT], gy[ROWS][VPT];  // xhat and dy*gamma
    float sums[2 * ROWS];
    float rs[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = r0 + r;
      sums[2 * r] = 0.f;
      sums[2 * r + 1] = 0.f;
      float mu = 0.f;
      rs[r] = 0.f;
      if (row < rows) {
        mu = __ldg(mean + row);
        rs[r] = __ldg(rstd + row);
      }
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        if (row < rows) {
          const size_t off = (size_t)row * D + (v * nthr + tid) * 4;
          const float4 xv = ld_nc_f4(x + off);
          const uint2 draw = ld_nc_v2(dy + off);
          const float2 d01 = unpack_bf16x2(draw.x), d23 = unpack_bf16x2(draw.y);
          float4 h;
          h.x = (xv.x - mu) * rs[r];
          h.y = (xv.y - mu) * rs[r];
          h.z = (xv.z - mu) * rs[r];
          h.w = (xv.w - mu) * rs[r];
          xh[r][v] = h;
          acc_dg[v].x += d01.x * h.x;
          acc_dg[v].y += d01.y * h.y;
          acc_dg[v].z += d23.x * h.z;
          acc_dg[v].w += d23.y * h.w;
          acc_db[v].x += d01.x;
          acc_db[v].y += d01.y;
          acc_db[v].z += d23.x;
          acc_db[v].w += d23.y;
          float4 t;
          t.x = d01.x * g[v].x;
          t.y = d01.y * g[v].y;
          t.z = d23.x * g[v].z;
          t.w = d23.y * g[v].w;
          gy[r][v] = t;
          sums[2 * r] += (t.x + t.y) + (t.z + t.w);
          sums[2 * r + 1] += (t.x * h.x + t.y * h.y) + (t.z * h.z + t.w * h.w);
        } else {
          xh[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          gy[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    block_sum<2 * ROWS>(sums, red);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = r0 + r;
      if (row >= rows) continue;
      const float c1 = sums[2 * r] * invD, c2 = sums[2 * r + 1] * invD;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const size_t off = (size_t)row * D + (v * nthr + tid) * 4;
        float4 o;
        o.x = rs[r] * (gy[r][v].x - c1 - xh[r][v].x * c2);
        o.y = rs[r] * (gy[r][v].y - c1 - xh[r][v].y * c2);
        o.z = rs[r] * (gy[r][v].z - c1 - xh[r][v].z * c2);
        o.w = rs[r] * (gy[r][v].w - c1 - xh[r][v].w * c2);
        if (dres_in != nullptr) {
          const float4 din = *reinterpret_cast<const float4*>(dres_in + off);
          o.x += din.x;
          o.y += din.y;
          o.z += din.z;
          o.w += din.w;
        }
        *reinterpret_cast<float4*>(dres_out + off) = o;
        const uint32_t p01 = pack_bf16x2(o.x, o.y), p23 = pack_bf16x2(o.z, o.w);
        if (dres_out_b != nullptr) *reinterpret_cast<uint2*>(dres_out_b + off) = make_uint2(p01, p23);
        const float2 r01 = unpack_bf16x2(p01), r23 = unpack_bf16x2(p23);
        acc_cs[v].x += r01.x;
        acc_cs[v].y += r01.y;
        acc_cs[v].z += r23.x;
        acc_cs[v].w += r23.y;
      }
    }
  }
  float* pbase = partials + (size_t)blockIdx.x * 3 * D;
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int col = (v * nthr + tid) * 4;
    *reinterpret_cast<float4*>(pbase + col) = acc_dg[v];
    *reinterpret_cast<float4*>(pbase + D + col) = acc_db[v];
    *reinterpret_cast<float4*>(pbase + 2 * D + col) = acc_cs[v];
  }
}

// out_k[col] (+)= sum_p partials[p][k][col]
__global__ void ln_bwd_finalize_kernel(const float* __restrict__ partials, int num_partials, int D, float* dgamma,
                                       float* dbeta, float* dcolsum, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over 3*D
  if (idx >= 3 * D) return;
  const int k = idx / D, col = idx - k * D;
  float* out = (k == 0) ? dgamma : (k == 1) ? dbeta : dcolsum;
  if (out == nullptr) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int pidx = 0;
  for (; pidx + 3 < num_partials; pidx += 4) {
    s0 += partials[((size_t)(pidx + 0) * 3 + k) * D + col];
    s1 += partials[((size_t)(pidx + 1) * 3 + k) * D + col];
    s2 += partials[((size_t)(pidx + 2) * 3 + k) * D + col];
    s3 += partials[((size_t)(pidx + 3) * 3 + k) * D + col];
  }
  for (; pidx < num_partials; ++pidx) s0 += partials[((size_t)pidx * 3 + k) * D + col];
  const float s = (s0 + s1) + (s2 + s3);
  out[col] = ((accumulate >> k) & 1) ? out[col] + s : s;
}

// choose VPT and block size: D = 4 * VPT * threads, threads multiple of 32, <= 512
static bool pick_config(int D, int* vpt, int* threads) {
  if (D % 128 != 0) return false;
  for (int v = 1; v <= 4; ++v) {
    if ((D / 4) % v != 0) continue;
    const int t = D / 4 / v;
    if (t % 32 == 0 && t <= 512 && t >= 32) {
      *vpt = v;
      *threads = t;
      return true;
    }
  }
  return false;
}

static int grid_for(int rows, int threads) {
  const int sms = num_sms();
  const int ctas_per_sm = threads <= 256 ? 4 : 2;
  int grid = sms * ctas_per_sm;
  const int need = (rows + ROWS - 1) / ROWS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  return grid;
}

}  // namespace ln
}  // namespace tae

extern "C" int tae_layernorm_fwd(const float* x, const float* gamma, const float* beta, tae_bf16* y, float* mean,
                                 float* rstd, int32_t rows, int32_t D, float eps, void* stream_) {
  using namespace tae;
  using namespace tae::ln;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(rows > 0 && D > 0, "tae_layernorm_fwd: rows/D must be positive");
  int vpt, threads;
  TAE_CHECK_SHAPE(pick_config(D, &vpt, &threads), "tae_layernorm_fwd: unsupported D=%d (need D %% 128 == 0, D <= 8192)", D);
  const int grid = grid_for(rows, threads);
  bf16* yy = reinterpret_cast<bf16*>(y);
  switch (vpt) {
    case 1: ln_fwd_kernel<1><<<grid, threads, 0, stream>>>(x, gamma, beta, yy, mean, rstd, rows, D, eps); break;
    case 2: ln_fwd_kernel<2><<<grid, threads, 0, stream>>>(x, gamma, beta, yy, mean, rstd, rows, D, eps); break;
    case 3: ln_fwd_kernel<3><<<grid, threads, 0, stream>>>(x, gamma, beta, yy, mean, rstd, rows, D, eps); break;
    default: ln_fwd_kernel<4><<<grid, threads, 0, stream>>>(x, gamma, beta, yy, mean, rstd, rows, D, eps); break;
  }
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_layernorm_bwd_num_partials(int32_t rows, int32_t D) {
  using namespace tae::ln;
  int vpt, threads;
  if (rows <= 0 || !pick_config(D, &vpt, &threads)) return TAE_ERR_SHAPE;
  return grid_for(rows, threads);
}

extern "C" int tae_layernorm_bwd(const tae_bf16* dy, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, const float* dres_in, float* dres_out, tae_bf16* dres_out_bf16,
                                 float* partials, int32_t rows, int32_t D, void* stream_) {
  using namespace tae;
  using namespace tae::ln;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(rows > 0 && D > 0, "tae_layernorm_bwd: rows/D must be positive");
  TAE_CHECK_SHAPE(dres_out != nullptr && partials != nullptr, "tae_layernorm_bwd: dres_out and partials are required");
  int vpt, threads;
  TAE_CHECK_SHAPE(pick_config(D, &vpt, &threads), "tae_layernorm_bwd: unsupported D=%d", D);
  const int grid = grid_for(rows, threads);
  const bf16* dyy = reinterpret_cast<const bf16*>(dy);
  bf16* ob = reinterpret_cast<bf16*>(dres_out_bf16);
  switch (vpt) {
    case 1: ln_bwd_kernel<1><<<grid, threads, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, D); break;
    case 2: ln_bwd_kernel<2><<<grid, threads, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, D); break;
    case 3: ln_bwd_kernel<3><<<grid, threads, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, D); break;
    default: ln_bwd_kernel<4><<<grid, threads, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, D); break;
  }
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_layernorm_bwd_finalize(const float* partials, int32_t num_partials, int32_t D, float* dgamma,
                                          float* dbeta, float* dcolsum, int32_t accumulate, void* stream_) {
  using namespace tae;
  using namespace tae::ln;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(num_partials > 0 && D > 0 && partials != nullptr, "tae_layernorm_bwd_finalize: bad arguments");
  const int total = 3 * D;
  ln_bwd_finalize_kernel<<<(total + 255) / 256, 256, 0, stream>>>(partials, num_partials, D, dgamma, dbeta, dcolsum,
                                                                   accumulate);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
