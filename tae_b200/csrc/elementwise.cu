// elementwise.cu — the HBM-bound kernels around the GEMMs: patch index maps, fused loss, reductions.
//
//   tae_im2col_bf16   PatchEmbed's strided conv as an im2col gather + bf16 cast          (tae.py:43,50-52)
//   tae_patchify / tae_unpatchify   the einsum permutes of TAE.patchify/unpatchify       (tae.py:196-222)
//   tae_mse_loss      forward_loss with patchify folded into the indexing, + dL/dpred     (tae.py:256-265)
//   tae_colsum_bf16   bias gradients (sum over tokens)
//   tae_batch_sum_f32 pos_embed gradients (sum over the batch)
// All index maps are pure integer arithmetic and bit-exact with the reference's permutes.
#include "common.cuh"

namespace tae {
namespace ew {

// ---------------------------------------------------------------------------------------------
// im2col: one thread converts 8 consecutive pixels of one image row (32 B in, 16 B out)
// ---------------------------------------------------------------------------------------------
__global__ void im2col_kernel(const float* __restrict__ imgs, bf16* __restrict__ cols, int B, int S, int p) {
  const int g = S / p;
  const int xg = S / 8;  // 8-pixel groups per image row
  const size_t total = (size_t)B * 3 * S * xg;
  const int Kp = 3 * p * p;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xq = (int)(idx % xg);
    size_t r = idx / xg;
    const int y = (int)(r % S);
    r /= S;
    const int c = (int)(r % 3);
    const int b = (int)(r / 3);
    const int x = xq * 8;
    const float* src = imgs + (((size_t)b * 3 + c) * S + y) * S + x;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    const int h = y / p, i = y - h * p, w = x / p, j = x - w * p;
    bf16* dst = cols + ((size_t)b * g * g + (size_t)h * g + w) * Kp + (size_t)c * p * p + i * p + j;
    uint4 o;
    o.x = pack_bf16x2(v0.x, v0.y);
    o.y = pack_bf16x2(v0.z, v0.w);
    o.z = pack_bf16x2(v1.x, v1.y);
    o.w = pack_bf16x2(v1.z, v1.w);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// patchify / unpatchify: out[b, h*g+w, (i*p+j)*3+c] <-> imgs[b, c, h*p+i, w*p+j]
// ---------------------------------------------------------------------------------------------
template <typename T, bool kUnpatchify>
__global__ void patch_permute_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int S, int p, int C) {
  const int g = S / p;
  const size_t per_img = (size_t)C * S * S;
  const size_t total = (size_t)B * per_img;
  const int Kp = C * p * p;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    // idx enumerates the patch layout [b][n][(i*p+j)*3+c]
    const int b = (int)(idx / per_img);
    const size_t r = idx - (size_t)b * per_img;
    const int n = (int)(r / Kp);
    const int e = (int)(r - (size_t)n * Kp);
    const int c = e % C;
    const int ij = e / C;
    const int i = ij / p, j = ij - i * p;
    const int h = n / g, w = n - h * g;
    const size_t img_off = (((size_t)b * C + c) * S + (size_t)h * p + i) * S + (size_t)w * p + j;
    if (kUnpatchify)
      dst[img_off] = src[idx];
    else
      dst[idx] = src[img_off];
  }
}

// ---------------------------------------------------------------------------------------------
// MSE loss + gradient, patchify folded into the index: one thread = 8 pixels x 3 channels = 24 pred elements
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mse_loss_kernel(const bf16* __restrict__ pred, const float* __restrict__ imgs, float* loss_accum,
                bf16* __restrict__ dpred, const float* __restrict__ grad_scale, int B, int S, int p) {
  const int g = S / p;
  const int jg = p / 8;  // 8-pixel groups per patch row
  const size_t total = (size_t)B * g * g * p * jg;
  const size_t numel = (size_t)B * 3 * S * S;
  const float inv_numel = 1.0f / (float)numel;
  const float gs = (grad_scale != nullptr ? __ldg(grad_scale) : 1.0f) * 2.0f * inv_numel;
  const int Kp = 3 * p * p;
  float lsum = 0.f;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int jq = (int)(idx % jg);
    size_t r = idx / jg;
    const int i = (int)(r % p);
    r /= p;
    const int n = (int)(r % (g * g));
    const int b = (int)(r / (g * g));
    const int h = n / g, w = n - h * g;
    const int j0 = jq * 8;
    const size_t poff = ((size_t)b * g * g + n) * Kp + (size_t)(i * p + j0) * 3;
    // L1-allocating loads: the three 16-byte pieces of a lane's 48 bytes share sectors with its neighbours' pieces
    const uint4* pp = reinterpret_cast<const uint4*>(pred + poff);
    const uint4 pr0 = __ldg(pp), pr1 = __ldg(pp + 1), pr2 = __ldg(pp + 2);
    const uint32_t pw[12] = {pr0.x, pr0.y, pr0.z, pr0.w, pr1.x, pr1.y, pr1.z, pr1.w, pr2.x, pr2.y, pr2.z, pr2.w};
    float pv[24];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float2 f = unpack_bf16x2(pw[k]);
      pv[2 * k] = f.x;
      pv[2 * k + 1] = f.y;
    }
    float tv[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* src = imgs + (((size_t)b * 3 + c) * S + (size_t)h * p + i) * S + (size_t)w * p + j0;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), bq = __ldg(reinterpret_cast<const float4*>(src) + 1);
      tv[c][0] = a.x; tv[c][1] = a.y; tv[c][2] = a.z; tv[c][3] = a.w;
      tv[c][4] = bq.x; tv[c][5] = bq.y; tv[c][6] = bq.z; tv[c][7] = bq.w;
    }
    float dv[24];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = pv[jj * 3 + c] - tv[c][jj];
        lsum = fmaf(d, d, lsum);
        dv[jj * 3 + c] = d * gs;
      }
    }
    if (dpred != nullptr) {
      uint32_t ow[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) ow[k] = pack_bf16x2(dv[2 * k], dv[2 * k + 1]);
      uint4* dst = reinterpret_cast<uint4*>(dpred + poff);
      dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      dst[2] = make_uint4(ow[8], ow[9], ow[10], ow[11]);
    }
  }
  __shared__ float red[8];
  lsum = warp_sum(lsum);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (warp == 0) {
    float s = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(loss_accum, s * inv_numel);
  }
}

// ---------------------------------------------------------------------------------------------
// column sums of a bf16 [M, N] matrix: stage 1 -> partial[chunk][N], stage 2 reduces the chunks
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_stage1(const bf16* __restrict__ x, int M, int N, int ldx, float* __restrict__ partial, int rows_per_chunk) {
  // block: 32 column groups (8 cols each = 256 cols) x 8 row lanes
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const int r_begin = blockIdx.y * rows_per_chunk;
  const int r_end = min(M, r_begin + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (col < N) {
    for (int r = r_begin + rl; r < r_end; r += 8) {
      const uint4 v = ld_nc_v4(x + (size_t)r * ldx + col);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
  }
  __shared__ float red[8][256 + 8];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cg * 8 + k] = acc[k];
  __syncthreads();
  const int c = threadIdx.x;  // 256 columns
  if (blockIdx.x * 256 + c < N) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][c];
    partial[(size_t)blockIdx.y * N + blockIdx.x * 256 + c] = s;
  }
}
__global__ void colsum_stage2(const float* __restrict__ partial, int chunks, int N, float* out, int accumulate) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * N + col];
  out[col] = accumulate ? out[col] + s : s;
}

static void colsum_config(int M, int N, int* chunks, int* rows_per_chunk) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int col_tiles = (N + 255) / 256;
  int want = (4 * sms + col_tiles - 1) / col_tiles;
  const int max_chunks = (M + 63) / 64;
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  int rpc = (M + want - 1) / want;
  rpc = (rpc + 7) / 8 * 8;
  *rows_per_chunk = rpc;
  *chunks = (M + rpc - 1) / rpc;
}

// out[n] (+)= sum_r x[r, n] for a (small) fp32 [R, N] matrix: block = 32 columns x 16 row-slices, 4 loads in flight
__global__ void __launch_bounds__(512) colsum_f32_kernel(const float* __restrict__ x, int R, int N, float* out, int accumulate) {
  __shared__ float red[16][33];
  const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (col < N) {
    int r = sl;
    for (; r + 48 < R; r += 64) {
      s0 += x[(size_t)r * N + col];
      s1 += x[(size_t)(r + 16) * N + col];
      s2 += x[(size_t)(r + 32) * N + col];
      s3 += x[(size_t)(r + 48) * N + col];
    }
    for (; r < R; r += 16) s0 += x[(size_t)r * N + col];
  }
  red[sl][cx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (sl == 0 && col < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += red[i][cx];
    out[col] = accumulate ? out[col] + s : s;
  }
}

// out[b, :] = mean_n x[b*N + n, :]   (global average pooling over tokens, tae.py:333)
__global__ void token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int N, int D) {
  const int d4 = blockIdx.x * blockDim.x + threadIdx.x;  // float4 column
  const int b = blockIdx.y;
  if (d4 * 4 >= D) return;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)b * N * D) + d4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int n = 0; n < N; ++n) {
    const float4 v = src[(size_t)n * (D / 4)];
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  const float inv = 1.0f / (float)N;
  reinterpret_cast<float4*>(out + (size_t)b * D)[d4] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}
// dx[b*N + n, :] = dy[b, :] / N
__global__ void token_mean_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int N, int D) {
  const size_t total = (size_t)B * N * (D / 4);
  const float inv = 1.0f / (float)N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int d4 = (int)(idx % (D / 4));
    const size_t row = idx / (D / 4);
    const int b = (int)(row / N);
    const float4 v = reinterpret_cast<const float4*>(dy + (size_t)b * D)[d4];
    reinterpret_cast<float4*>(dx)[idx] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}

// out[r, :] (+)= sum_b x[b*R + r, :]
__global__ void batch_sum_kernel(const float* __restrict__ x, int B, int R, int D, float* out, int accumulate) {
  const int d4 = D / 4;
  const size_t total = (size_t)R * d4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t off = idx * 4;  // element offset inside one [R, D] slab
  const size_t slab = (size_t)R * D;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  int b = 0;
  for (; b + 3 < B; b += 4) {
    const float4 v0 = ld_nc_f4(x + (size_t)(b + 0) * slab + off);
    const float4 v1 = ld_nc_f4(x + (size_t)(b + 1) * slab + off);
    const float4 v2 = ld_nc_f4(x + (size_t)(b + 2) * slab + off);
    const float4 v3 = ld_nc_f4(x + (size_t)(b + 3) * slab + off);
    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
    a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
    a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
  }
  for (; b < B; ++b) {
    const float4 v0 = ld_nc_f4(x + (size_t)b * slab + off);
    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
  }
  float4 s;
  s.x = (a0.x + a1.x) + (a2.x + a3.x);
  s.y = (a0.y + a1.y) + (a2.y + a3.y);
  s.z = (a0.z + a1.z) + (a2.z + a3.z);
  s.w = (a0.w + a1.w) + (a2.w + a3.w);
  float4* o = reinterpret_cast<float4*>(out + off);
  if (accumulate) {
    const float4 prev = *o;
    s.x += prev.x; s.y += prev.y; s.z += prev.z; s.w += prev.w;
  }
  *o = s;
}

static int stream_grid(size_t total_threads, int block) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  size_t blocks = (total_threads + block - 1) / block;
  const size_t cap = (size_t)sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace ew
}  // namespace tae

extern "C" int tae_im2col_bf16(const float* imgs, tae_bf16* cols, int32_t B, int32_t S, int32_t p, void* stream_) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && S > 0 && p > 0 && S % p == 0, "tae_im2col_bf16: need S %% p == 0 (S=%d p=%d)", S, p);
  TAE_CHECK_SHAPE(p % 8 == 0, "tae_im2col_bf16: patch size must be a multiple of 8 (p=%d)", p);
  const size_t total = (size_t)B * 3 * S * (S / 8);
  im2col_kernel<<<stream_grid(total, 256), 256, 0, stream>>>(imgs, reinterpret_cast<bf16*>(cols), B, S, p);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

static int patch_permute(const void* src, void* dst, int32_t B, int32_t S, int32_t p, int32_t elem_size, bool unpatchify,
                         void* stream_, int32_t C = 3) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && S > 0 && p > 0 && S % p == 0, "patchify/unpatchify: need S %% p == 0 (S=%d p=%d)", S, p);
  TAE_CHECK_SHAPE(elem_size == 2 || elem_size == 4, "patchify/unpatchify: elem_size must be 2 or 4");
  TAE_CHECK_SHAPE(C > 0, "patchify/unpatchify: channels must be positive");
  const size_t total = (size_t)B * C * S * S;
  const int grid = stream_grid(total, 256);
  if (elem_size == 4) {
    if (unpatchify)
      patch_permute_kernel<uint32_t, true><<<grid, 256, 0, stream>>>((const uint32_t*)src, (uint32_t*)dst, B, S, p, C);
    else
      patch_permute_kernel<uint32_t, false><<<grid, 256, 0, stream>>>((const uint32_t*)src, (uint32_t*)dst, B, S, p, C);
  } else {
    if (unpatchify)
      patch_permute_kernel<uint16_t, true><<<grid, 256, 0, stream>>>((const uint16_t*)src, (uint16_t*)dst, B, S, p, C);
    else
      patch_permute_kernel<uint16_t, false><<<grid, 256, 0, stream>>>((const uint16_t*)src, (uint16_t*)dst, B, S, p, C);
  }
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_patchify(const void* imgs, void* out, int32_t B, int32_t S, int32_t p, int32_t elem_size, void* stream) {
  return patch_permute(imgs, out, B, S, p, elem_size, false, stream);
}
extern "C" int tae_unpatchify(const void* x, void* imgs, int32_t B, int32_t S, int32_t p, int32_t elem_size, void* stream) {
  return patch_permute(x, imgs, B, S, p, elem_size, true, stream);
}

/* C-channel variants: VITForSegmentation.unpatchify (tae.py:391-403) and its adjoint */
extern "C" int tae_patchify_c(const void* imgs, void* out, int32_t B, int32_t S, int32_t p, int32_t C, int32_t elem_size,
                              void* stream) {
  return patch_permute(imgs, out, B, S, p, elem_size, false, stream, C);
}
extern "C" int tae_unpatchify_c(const void* x, void* imgs, int32_t B, int32_t S, int32_t p, int32_t C, int32_t elem_size,
                                void* stream) {
  return patch_permute(x, imgs, B, S, p, elem_size, true, stream, C);
}

extern "C" int tae_mse_loss(const tae_bf16* pred, const float* imgs, float* loss_accum, tae_bf16* dpred,
                            const float* grad_scale, int32_t B, int32_t S, int32_t p, void* stream_) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && S > 0 && p > 0 && S % p == 0 && p % 8 == 0, "tae_mse_loss: need S %% p == 0 and p %% 8 == 0 (S=%d p=%d)", S, p);
  TAE_CHECK_SHAPE(loss_accum != nullptr, "tae_mse_loss: loss_accum is NULL");
  const size_t total = (size_t)B * 3 * S * S / 24;
  int grid = stream_grid(total, 256);
  const int sms = num_sms() > 0 ? num_sms() : 148;
  if (grid > sms * 4) grid = sms * 4;  // bound the number of loss atomics
  mse_loss_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(pred), imgs, loss_accum,
                                            reinterpret_cast<bf16*>(dpred), grad_scale, B, S, p);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" size_t tae_colsum_workspace_floats(int32_t M, int32_t N) {
  if (M <= 0 || N <= 0) return 0;
  int chunks, rpc;
  tae::ew::colsum_config(M, N, &chunks, &rpc);
  return (size_t)chunks * (size_t)N;
}

extern "C" int tae_colsum_bf16(const tae_bf16* x, int32_t M, int32_t N, int32_t ldx, float* out, int32_t accumulate,
                               float* workspace, void* stream_) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(M > 0 && N > 0 && N % 8 == 0 && ldx % 8 == 0 && ldx >= N, "tae_colsum_bf16: bad shape M=%d N=%d ldx=%d", M, N, ldx);
  TAE_CHECK_SHAPE(workspace != nullptr && out != nullptr, "tae_colsum_bf16: NULL workspace/out");
  int chunks, rpc;
  colsum_config(M, N, &chunks, &rpc);
  dim3 grid((N + 255) / 256, chunks);
  colsum_stage1<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), M, N, ldx, workspace, rpc);
  TAE_CHECK_LAUNCH();
  colsum_stage2<<<(N + 255) / 256, 256, 0, stream>>>(workspace, chunks, N, out, accumulate);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_colsum_f32(const float* x, int32_t R, int32_t N, float* out, int32_t accumulate, void* stream_) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(R > 0 && N > 0 && x != nullptr && out != nullptr, "tae_colsum_f32: bad arguments");
  colsum_f32_kernel<<<(N + 31) / 32, 512, 0, stream>>>(x, R, N, out, accumulate);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_token_mean_f32(const float* x, float* out, int32_t B, int32_t N, int32_t D, void* stream_) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && N > 0 && D > 0 && D % 4 == 0 && x && out, "tae_token_mean_f32: bad shape B=%d N=%d D=%d", B, N, D);
  dim3 grid((D / 4 + 127) / 128, B);
  token_mean_kernel<<<grid, 128, 0, stream>>>(x, out, B, N, D);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
extern "C" int tae_token_mean_bwd_f32(const float* dy, float* dx, int32_t B, int32_t N, int32_t D, void* stream_) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && N > 0 && D > 0 && D % 4 == 0 && dy && dx, "tae_token_mean_bwd_f32: bad shape B=%d N=%d D=%d", B, N, D);
  token_mean_bwd_kernel<<<stream_grid((size_t)B * N * (D / 4), 256), 256, 0, stream>>>(dy, dx, B, N, D);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_batch_sum_f32(const float* x, int32_t B, int32_t R, int32_t D, float* out, int32_t accumulate,
                                 void* stream_) {
  using namespace tae;
  using namespace tae::ew;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(B > 0 && R > 0 && D > 0 && D % 4 == 0, "tae_batch_sum_f32: bad shape B=%d R=%d D=%d", B, R, D);
  const size_t total = (size_t)R * (D / 4);
  batch_sum_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(x, B, R, D, out, accumulate);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
