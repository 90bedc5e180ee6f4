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

constexpr int FWD_ROWS = 4;   // rows per CTA iteration (forward)
constexpr int BWD_ROWS = 2;   // rows per CTA iteration (backward: more live state per row)
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

// CTA-wide (mean, M2) of NR rows in ONE barrier round: every thread enters with the mean / sum of squared deviations
// of its own `cnt0` elements; equal-count pairwise merges (Chan et al.) through the shuffle tree, then a sequential
// merge over the warps.  Numerically a two-pass variance, at the synchronisation cost of a single sum.
template <int NR>
__device__ __forceinline__ void block_mean_m2(float (&mean)[NR], float (&m2)[NR], float cnt0, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float cnt = cnt0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float mb = __shfl_xor_sync(0xffffffffu, mean[r], o);
      const float qb = __shfl_xor_sync(0xffffffffu, m2[r], o);
      const float d = mb - mean[r];
      mean[r] = 0.5f * (mean[r] + mb);
      m2[r] = m2[r] + qb + d * d * (0.5f * cnt);
    }
    cnt *= 2.0f;
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      red[(2 * r) * MAX_WARPS + warp] = mean[r];
      red[(2 * r + 1) * MAX_WARPS + warp] = m2[r];
    }
  }
  __syncthreads();
  // sequential equal-count merge over the warps: after w groups, n = w*cnt, so cnt/(n+cnt) = 1/(w+1)
  for (int w = 1; w < nwarps; ++w) {
    const float inv = __frcp_rn((float)(w + 1));
    const float wgt = cnt * (float)w * inv;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float mu = (w == 1) ? red[(2 * r) * MAX_WARPS] : mean[r];
      const float q = (w == 1) ? red[(2 * r + 1) * MAX_WARPS] : m2[r];
      const float mw = red[(2 * r) * MAX_WARPS + w], qw = red[(2 * r + 1) * MAX_WARPS + w];
      const float d = mw - mu;
      mean[r] = fmaf(d, inv, mu);
      m2[r] = q + qw + d * d * wgt;
    }
  }
  if (nwarps == 1) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      mean[r] = red[(2 * r) * MAX_WARPS];
      m2[r] = red[(2 * r + 1) * MAX_WARPS];
    }
  }
}

template <int VPT>
__global__ void __launch_bounds__(512)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D,
              float eps) {
  constexpr int ROWS = FWD_ROWS;
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
    float mu[ROWS], q[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = r0 + r;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        if (row < rows) {
          xv[r][v] = ld_nc_f4(x + (size_t)row * D + (v * nthr + tid) * 4);
        } else {
          xv[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < VPT; ++v) s += (xv[r][v].x + xv[r][v].y) + (xv[r][v].z + xv[r][v].w);
      mu[r] = s * (1.0f / (4 * VPT));
      q[r] = 0.f;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const float dx = xv[r][v].x - mu[r], dy = xv[r][v].y - mu[r], dz = xv[r][v].z - mu[r], dw = xv[r][v].w - mu[r];
        q[r] += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    block_mean_m2<ROWS>(mu, q, (float)(4 * VPT), red);
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

// Forward, warp-per-row variant (D = NV*128 <= 4096): the whole row lives in one warp's registers, statistics are
// two shuffle reductions (mean, then centred sum of squares) and there is no block-level barrier at all; every lane
// has NV independent 128-bit loads in flight.  gamma/beta are re-read through L1 (they are 2*D*4 bytes, L1-resident).
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_warp_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                   float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float invD = 1.0f / (float)D;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const float* xr = x + (size_t)row * D;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = ld_nc_f4(xr + (i * 32 + lane) * 4);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mu = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rs = rsqrtf(warp_sum(q) * invD + eps);
    if (lane == 0) {
      mean_out[row] = mu;
      rstd_out[row] = rs;
    }
    bf16* yr = y + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
      uint2 o;
      o.x = pack_bf16x2((v[i].x - mu) * rs * g.x + b.x, (v[i].y - mu) * rs * g.y + b.y);
      o.y = pack_bf16x2((v[i].z - mu) * rs * g.z + b.z, (v[i].w - mu) * rs * g.w + b.w);
      *reinterpret_cast<uint2*>(yr + col) = o;
    }
  }
}

template <int NV>
static void launch_fwd_warp(const float* x, const float* gamma, const float* beta, bf16* y, float* mean, float* rstd,
                            int rows, float eps, cudaStream_t stream) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int grid = (rows + 7) / 8;
  if (grid > sms * 8) grid = sms * 8;
  ln_fwd_warp_kernel<NV><<<grid, 256, 0, stream>>>(x, gamma, beta, y, mean, rstd, rows, eps);
}

// Backward.  partials layout: [gridDim.x][3][D]  (0: dgamma, 1: dbeta, 2: colsum(bf16(dres_out)))
template <int VPT>
__global__ void __launch_bounds__(512, (VPT == 1 ? 2 : 1))
ln_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* dres_in,
              float* dres_out, bf16* __restrict__ dres_out_b, float* __restrict__ partials, int rows, int D) {
  constexpr int ROWS = BWD_ROWS;
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
    float4 xh[ROWS][VPT], gy[ROWS][VPT], din[ROWS][VPT];  // xhat, dy*gamma, incoming residual gradient
    float sums[2 * ROWS];
    float rs[ROWS];
    // issue every global load of this iteration up front (x, dy AND the residual gradient)
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = r0 + r;
      float mu = 0.f;
      rs[r] = 0.f;
      if (row < rows) {
        mu = __ldg(mean + row);
        rs[r] = __ldg(rstd + row);
      }
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        xh[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        gy[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        din[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) {
          const size_t off = (size_t)row * D + (v * nthr + tid) * 4;
          const float4 xv = ld_nc_f4(x + off);
          const uint2 draw = ld_nc_v2(dy + off);
          if (dres_in != nullptr) din[r][v] = *reinterpret_cast<const float4*>(dres_in + off);
          const float2 d01 = unpack_bf16x2(draw.x), d23 = unpack_bf16x2(draw.y);
          xh[r][v] = make_float4((xv.x - mu) * rs[r], (xv.y - mu) * rs[r], (xv.z - mu) * rs[r], (xv.w - mu) * rs[r]);
          gy[r][v] = make_float4(d01.x, d01.y, d23.x, d23.y);  // dy for now
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      sums[2 * r] = 0.f;
      sums[2 * r + 1] = 0.f;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const float4 d = gy[r][v], h = xh[r][v];
        acc_dg[v].x += d.x * h.x;
        acc_dg[v].y += d.y * h.y;
        acc_dg[v].z += d.z * h.z;
        acc_dg[v].w += d.w * h.w;
        acc_db[v].x += d.x;
        acc_db[v].y += d.y;
        acc_db[v].z += d.z;
        acc_db[v].w += d.w;
        const float4 t = make_float4(d.x * g[v].x, d.y * g[v].y, d.z * g[v].z, d.w * g[v].w);
        gy[r][v] = t;
        sums[2 * r] += (t.x + t.y) + (t.z + t.w);
        sums[2 * r + 1] += (t.x * h.x + t.y * h.y) + (t.z * h.z + t.w * h.w);
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
        o.x = rs[r] * (gy[r][v].x - c1 - xh[r][v].x * c2) + din[r][v].x;
        o.y = rs[r] * (gy[r][v].y - c1 - xh[r][v].y * c2) + din[r][v].y;
        o.z = rs[r] * (gy[r][v].z - c1 - xh[r][v].z * c2) + din[r][v].z;
        o.w = rs[r] * (gy[r][v].w - c1 - xh[r][v].w * c2) + din[r][v].w;
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

// Backward, warp-per-row variant (D = NV*128 <= 1024).  A warp owns whole rows, so the two row statistics are shuffle
// reductions and the kernel has no block-level barrier; a lane owns the same 4*NV columns for every row its warp
// visits, so dgamma / dbeta / colsum accumulate in registers.  One partial row per WARP.
template <int NV>
__global__ void __launch_bounds__(256, 1)
ln_bwd_warp_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const float* dres_in,
                   float* dres_out, bf16* __restrict__ dres_out_b, float* __restrict__ partials, int rows, int* sched) {
  constexpr int D = NV * 128;
  constexpr int CHUNK = 1;  // rows per dynamically scheduled work item (1 keeps concurrent warps on consecutive rows)
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int gw = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int nw = gridDim.x * wpb;
  float4 g[NV], acc_dg[NV], acc_db[NV], acc_cs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4));
    acc_dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invD = 1.0f / (float)D;
  // Work list: static row striding, or (sched != NULL) chunks of CHUNK rows drawn from a global counter one ahead, so a
  // CTA that starts late (its SM held by another kernel, e.g. NCCL) does not leave its whole share for the end.
  int next_chunk = 0;
  if (sched != nullptr) {
    if (lane == 0) next_chunk = atomicAdd(sched, 1);
    next_chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
  }
#pragma unroll 1
  for (int it = 0;; ++it) {
    int row_begin, row_end;
    if (sched != nullptr) {
      row_begin = next_chunk * CHUNK;
      if (row_begin >= rows) break;
      row_end = min(rows, row_begin + CHUNK);
      if (lane == 0) next_chunk = atomicAdd(sched, 1);
      next_chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
    } else {
      row_begin = gw + it * nw;
      if (row_begin >= rows) break;
      row_end = row_begin + 1;
    }
#pragma unroll 1
  for (int row = row_begin; row < row_end; ++row) {
    const size_t base = (size_t)row * D;
    float4 xh[NV], gy[NV], din[NV];
    uint2 draw[NV];
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const size_t off = base + (i * 32 + lane) * 4;
      xh[i] = ld_nc_f4(x + off);
      draw[i] = ld_nc_v2(dy + off);
      din[i] = dres_in != nullptr ? *reinterpret_cast<const float4*>(dres_in + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float2 d01 = unpack_bf16x2(draw[i].x), d23 = unpack_bf16x2(draw[i].y);
      const float4 d = make_float4(d01.x, d01.y, d23.x, d23.y);
      const float4 h = make_float4((xh[i].x - mu) * rs, (xh[i].y - mu) * rs, (xh[i].z - mu) * rs, (xh[i].w - mu) * rs);
      xh[i] = h;
      acc_dg[i].x += d.x * h.x;
      acc_dg[i].y += d.y * h.y;
      acc_dg[i].z += d.z * h.z;
      acc_dg[i].w += d.w * h.w;
      acc_db[i].x += d.x;
      acc_db[i].y += d.y;
      acc_db[i].z += d.z;
      acc_db[i].w += d.w;
      const float4 t = make_float4(d.x * g[i].x, d.y * g[i].y, d.z * g[i].z, d.w * g[i].w);
      gy[i] = t;
      s1 += (t.x + t.y) + (t.z + t.w);
      s2 += (t.x * h.x + t.y * h.y) + (t.z * h.z + t.w * h.w);
    }
    const float c1 = warp_sum(s1) * invD, c2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const size_t off = base + (i * 32 + lane) * 4;
      float4 o;
      o.x = rs * (gy[i].x - c1 - xh[i].x * c2) + din[i].x;
      o.y = rs * (gy[i].y - c1 - xh[i].y * c2) + din[i].y;
      o.z = rs * (gy[i].z - c1 - xh[i].z * c2) + din[i].z;
      o.w = rs * (gy[i].w - c1 - xh[i].w * c2) + din[i].w;
      *reinterpret_cast<float4*>(dres_out + off) = o;
      const uint32_t p01 = pack_bf16x2(o.x, o.y), p23 = pack_bf16x2(o.z, o.w);
      if (dres_out_b != nullptr) *reinterpret_cast<uint2*>(dres_out_b + off) = make_uint2(p01, p23);
      const float2 r01 = unpack_bf16x2(p01), r23 = unpack_bf16x2(p23);
      acc_cs[i].x += r01.x;
      acc_cs[i].y += r01.y;
      acc_cs[i].z += r23.x;
      acc_cs[i].w += r23.y;
    }
  }
  }
  if (sched != nullptr && lane == 0) {
    // every warp has drawn its last (out-of-range) chunk before it counts itself done: the last one re-arms the slot
    if (atomicAdd(sched + 1, 1) == nw - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
  float* pbase = partials + (size_t)gw * 3 * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(pbase + col) = acc_dg[i];
    *reinterpret_cast<float4*>(pbase + D + col) = acc_db[i];
    *reinterpret_cast<float4*>(pbase + 2 * D + col) = acc_cs[i];
  }
}

// Backward, ROW-GROUP variant for the wide residual streams (D = W*512: 2048 for patch32, 2560 for patch64 / patch128).
// The warp-per-row kernel above would need 16*NV registers of per-column accumulators (NV = 16, 20); the block-per-row
// kernel synchronises the whole CTA twice per row pair and keeps one CTA per SM (0.59 / 0.33 / 0.21 of the HBM peak at
// 16384 / 4096 / 1024 rows).  Here W warps share a row (a lane owns 4 column quads, interleaved so that every warp
// reads 512 contiguous bytes per access), R such groups per CTA walk rows independently, and the only synchronisation is
// ONE named barrier per row among the group's W warps (the two row sums travel through double-buffered shared-memory
// slots, together with the next row index when rows are drawn dynamically).  The R groups fold their dgamma / dbeta /
// colsum accumulators through shared memory at the end: one partial row per CTA, so small-M launches (1024 rows) no
// longer write and re-read more partial-sum bytes than the finalize pass is worth.
__device__ __forceinline__ void group_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int W, int R>
__global__ void __launch_bounds__(32 * W * R, 1)
ln_bwd_group_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* dres_in,
                    float* dres_out, bf16* __restrict__ dres_out_b, float* __restrict__ partials, int rows, int* sched) {
  constexpr int NV = 4;
  constexpr int D = W * 512;
  __shared__ float s_stat[R][2][W][2];
  __shared__ int s_next[R][2];
  __shared__ __align__(16) float s_acc[3 * D];
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int rg = wic / W, w = wic - rg * W;
  const int group = blockIdx.x * R + rg, ngroups = gridDim.x * R;
  float4 g[NV], acc_dg[NV], acc_db[NV], acc_cs[NV];
  int col[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    col[i] = ((i * W + w) * 32 + lane) * 4;
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma + col[i]));
    acc_dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invD = 1.0f / (float)D;
  const bool dynamic = sched != nullptr;
  int row = group;
  if (dynamic) {
    if (w == 0 && lane == 0) s_next[rg][1] = atomicAdd(sched, 1);
    group_bar_sync(1 + rg, W * 32);
    row = s_next[rg][1];
  }
  int buf = 0;
#pragma unroll 1
  while (row < rows) {
    const size_t base = (size_t)row * D;
    float4 xh[NV], gy[NV], din[NV];
    uint2 draw[NV];
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const size_t off = base + col[i];
      xh[i] = ld_nc_f4(x + off);
      draw[i] = ld_nc_v2(dy + off);
      din[i] = dres_in != nullptr ? *reinterpret_cast<const float4*>(dres_in + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // the group's leader draws the NEXT row while this row's loads are in flight
    if (dynamic && w == 0 && lane == 0) s_next[rg][buf] = atomicAdd(sched, 1);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float2 d01 = unpack_bf16x2(draw[i].x), d23 = unpack_bf16x2(draw[i].y);
      const float4 d = make_float4(d01.x, d01.y, d23.x, d23.y);
      const float4 h = make_float4((xh[i].x - mu) * rs, (xh[i].y - mu) * rs, (xh[i].z - mu) * rs, (xh[i].w - mu) * rs);
      xh[i] = h;
      acc_dg[i].x += d.x * h.x;
      acc_dg[i].y += d.y * h.y;
      acc_dg[i].z += d.z * h.z;
      acc_dg[i].w += d.w * h.w;
      acc_db[i].x += d.x;
      acc_db[i].y += d.y;
      acc_db[i].z += d.z;
      acc_db[i].w += d.w;
      const float4 t = make_float4(d.x * g[i].x, d.y * g[i].y, d.z * g[i].z, d.w * g[i].w);
      gy[i] = t;
      s1 += (t.x + t.y) + (t.z + t.w);
      s2 += (t.x * h.x + t.y * h.y) + (t.z * h.z + t.w * h.w);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      s_stat[rg][buf][w][0] = s1;
      s_stat[rg][buf][w][1] = s2;
    }
    group_bar_sync(1 + rg, W * 32);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < W; ++k) {  // same order in every warp: all lanes of the row agree bit for bit
      c1 += s_stat[rg][buf][k][0];
      c2 += s_stat[rg][buf][k][1];
    }
    c1 *= invD;
    c2 *= invD;
    const int next_row = dynamic ? s_next[rg][buf] : row + ngroups;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const size_t off = base + col[i];
      float4 o;
      o.x = rs * (gy[i].x - c1 - xh[i].x * c2) + din[i].x;
      o.y = rs * (gy[i].y - c1 - xh[i].y * c2) + din[i].y;
      o.z = rs * (gy[i].z - c1 - xh[i].z * c2) + din[i].z;
      o.w = rs * (gy[i].w - c1 - xh[i].w * c2) + din[i].w;
      *reinterpret_cast<float4*>(dres_out + off) = o;
      const uint32_t p01 = pack_bf16x2(o.x, o.y), p23 = pack_bf16x2(o.z, o.w);
      if (dres_out_b != nullptr) *reinterpret_cast<uint2*>(dres_out_b + off) = make_uint2(p01, p23);
      const float2 r01 = unpack_bf16x2(p01), r23 = unpack_bf16x2(p23);
      acc_cs[i].x += r01.x;
      acc_cs[i].y += r01.y;
      acc_cs[i].z += r23.x;
      acc_cs[i].w += r23.y;
    }
    row = next_row;
    buf ^= 1;
  }
  if (dynamic && w == 0 && lane == 0) {
    // every group has drawn its last (out-of-range) row before it counts itself done: the last one re-arms the slot
    if (atomicAdd(sched + 1, 1) == ngroups - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
  // fold the R groups' column accumulators through shared memory: one partial row [3][D] per CTA
#pragma unroll 1
  for (int r = 0; r < R; ++r) {
    if (rg == r) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4* a0 = reinterpret_cast<float4*>(s_acc + col[i]);
        float4* a1 = reinterpret_cast<float4*>(s_acc + D + col[i]);
        float4* a2 = reinterpret_cast<float4*>(s_acc + 2 * D + col[i]);
        if (r == 0) {
          *a0 = acc_dg[i];
          *a1 = acc_db[i];
          *a2 = acc_cs[i];
        } else {
          float4 v = *a0;
          *a0 = make_float4(v.x + acc_dg[i].x, v.y + acc_dg[i].y, v.z + acc_dg[i].z, v.w + acc_dg[i].w);
          v = *a1;
          *a1 = make_float4(v.x + acc_db[i].x, v.y + acc_db[i].y, v.z + acc_db[i].z, v.w + acc_db[i].w);
          v = *a2;
          *a2 = make_float4(v.x + acc_cs[i].x, v.y + acc_cs[i].y, v.z + acc_cs[i].z, v.w + acc_cs[i].w);
        }
      }
    }
    __syncthreads();
  }
  float4* pdst = reinterpret_cast<float4*>(partials + (size_t)blockIdx.x * 3 * D);
  for (int i = threadIdx.x; i < 3 * D / 4; i += blockDim.x) pdst[i] = reinterpret_cast<const float4*>(s_acc)[i];
}

static bool bwd_group_ok(int D) { return D == 2048 || D == 2560; }
static int bwd_group_grid(int rows, int R) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int grid = (rows + R - 1) / R;
  if (grid > sms) grid = sms;
  return grid < 1 ? 1 : grid;
}
constexpr int GROUP_R_2048 = 3, GROUP_R_2560 = 2;  // row groups per CTA: 384 / 320 threads, 60 / 51 KB of loads in flight

// warp-per-row backward: which widths, and how many partial rows (= warps) it produces
static bool bwd_warp_ok(int D) { return D % 128 == 0 && D / 128 <= 8 && (D / 128 == 1 || D / 128 == 2 || D / 128 == 6 || D / 128 == 8); }
static int bwd_warp_grid(int rows) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int grid = (rows + 7) / 8;
  if (grid > sms) grid = sms;
  return grid < 1 ? 1 : grid;
}

// out_k[col] (+)= sum_p partials[p][k][col].  Block = 32 columns x 32 partial-slices (coalesced 128-byte rows),
// grid = 3*D/32 blocks; every thread keeps 4 independent loads in flight, so the few hundred partial rows cost
// ~5 dependent round trips instead of ~40.
__global__ void __launch_bounds__(1024)
ln_bwd_finalize_kernel(const float* __restrict__ partials, int num_partials, int D, float* dgamma, float* dbeta,
                       float* dcolsum, int accumulate) {
  __shared__ float red[32][33];
  const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + cx;  // over 3*D
  const int k = idx / D, col = idx - k * D;  // D % 32 == 0: a block never straddles two outputs
  float* out = (k == 0) ? dgamma : (k == 1) ? dbeta : dcolsum;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (out != nullptr) {
    const float* src = partials + (size_t)k * D + col;
    const size_t stride = (size_t)3 * D;
    int p = sl;
    for (; p + 96 < num_partials; p += 128) {
      s0 += src[(size_t)p * stride];
      s1 += src[(size_t)(p + 32) * stride];
      s2 += src[(size_t)(p + 64) * stride];
      s3 += src[(size_t)(p + 96) * stride];
    }
    for (; p < num_partials; p += 32) s0 += src[(size_t)p * stride];
  }
  red[sl][cx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (sl == 0 && out != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red[i][cx];
    out[col] = ((accumulate >> k) & 1) ? out[col] + s : s;
  }
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

static int grid_for(int rows, int threads, bool bwd) {
  const int sms = num_sms();
  // forward: fill the SM's 2048 thread slots; backward: 64-register kernels -> 1024 threads per SM
  const int ctas_per_sm = (bwd ? 1024 : 2048) / threads > 0 ? (bwd ? 1024 : 2048) / threads : 1;
  int grid = sms * ctas_per_sm;
  const int rows_per_it = bwd ? BWD_ROWS : FWD_ROWS;
  const int need = (rows + rows_per_it - 1) / rows_per_it;
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
  const int grid = grid_for(rows, threads, false);
  bf16* yy = reinterpret_cast<bf16*>(y);
  bool done = true;
  switch (D / 128) {  // warp-per-row kernel for the widths the model zoo uses
    case 1: launch_fwd_warp<1>(x, gamma, beta, yy, mean, rstd, rows, eps, stream); break;
    case 2: launch_fwd_warp<2>(x, gamma, beta, yy, mean, rstd, rows, eps, stream); break;
    case 5: launch_fwd_warp<5>(x, gamma, beta, yy, mean, rstd, rows, eps, stream); break;
    case 6: launch_fwd_warp<6>(x, gamma, beta, yy, mean, rstd, rows, eps, stream); break;
    case 8: launch_fwd_warp<8>(x, gamma, beta, yy, mean, rstd, rows, eps, stream); break;
    case 16: launch_fwd_warp<16>(x, gamma, beta, yy, mean, rstd, rows, eps, stream); break;
    case 20: launch_fwd_warp<20>(x, gamma, beta, yy, mean, rstd, rows, eps, stream); break;
    default: done = false;
  }
  if (done) {
    TAE_CHECK_LAUNCH();
    return TAE_OK;
  }
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
  if (bwd_warp_ok(D)) return bwd_warp_grid(rows) * 8;
  if (bwd_group_ok(D)) return bwd_group_grid(rows, D == 2048 ? GROUP_R_2048 : GROUP_R_2560);
  return grid_for(rows, threads, true);
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
  const int grid = grid_for(rows, threads, true);
  const bf16* dyy = reinterpret_cast<const bf16*>(dy);
  bf16* ob = reinterpret_cast<bf16*>(dres_out_bf16);
  if (bwd_warp_ok(D)) {
    const int wg = bwd_warp_grid(rows);
    int* sched = rows > wg * 8 ? sched_counter_slot() : nullptr;  // more rows than warps: dynamic row chunks
    switch (D / 128) {
      case 1: ln_bwd_warp_kernel<1><<<wg, 256, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, sched); break;
      case 2: ln_bwd_warp_kernel<2><<<wg, 256, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, sched); break;
      case 6: ln_bwd_warp_kernel<6><<<wg, 256, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, sched); break;
      default: ln_bwd_warp_kernel<8><<<wg, 256, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob, partials, rows, sched); break;
    }
    TAE_CHECK_LAUNCH();
    return TAE_OK;
  }
  if (bwd_group_ok(D)) {
    const int R = D == 2048 ? GROUP_R_2048 : GROUP_R_2560;
    const int gg = bwd_group_grid(rows, R);
    int* sched = rows > gg * R ? sched_counter_slot() : nullptr;  // more rows than groups: dynamic rows (if enabled)
    if (D == 2048)
      ln_bwd_group_kernel<4, GROUP_R_2048><<<gg, 32 * 4 * GROUP_R_2048, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob,
                                                                                   partials, rows, sched);
    else
      ln_bwd_group_kernel<5, GROUP_R_2560><<<gg, 32 * 5 * GROUP_R_2560, 0, stream>>>(dyy, x, mean, rstd, gamma, dres_in, dres_out, ob,
                                                                                   partials, rows, sched);
    TAE_CHECK_LAUNCH();
    return TAE_OK;
  }
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
  ln_bwd_finalize_kernel<<<total / 32, 1024, 0, stream>>>(partials, num_partials, D, dgamma, dbeta, dcolsum, accumulate);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
