// attention.cu — short-sequence attention, whole sequence resident in shared memory.
//
// Replaces F.scaled_dot_product_attention(q, k, v) (tae.py:78; non-causal, no mask, scale 1/sqrt(hd)) together
// with the qkv un-bind permute (tae.py:74-75) and the head merge (tae.py:80): the kernels index the packed
// [B*N, 3*H*hd] qkv buffer directly and write the merged [B*N, H*hd] layout, so neither permute is materialised.
//
// Code paths:
//   * N = 256, hd = 64 (patch16): the tcgen05 kernels of attention_sm100.cu;
//   * N = 64, hd = 64 (patch32): persistent CTAs, one warp per 16 query rows / 16 keys, bf16 mma.sync.m16n8k16 with fp32
//     accumulation, exp2-domain online softmax with quad shuffles.  Backward recomputes P from the saved log-sum-exp:
//     dK/dV with one warp per 16 keys, dS^T handed through shared memory, then dQ with one warp per 16 queries — no
//     atomics, deterministic;
//   * N <= 16 (patch64 / patch128, hd 32/64/80): one warp per (image, head), mma.sync on a padded 16x16 score tile;
//   * generic path (any N <= 128, hd <= 128; the 4- and 16-token grids of patch128 / patch64 with hd = 80):
//     one warp per (image, head), fp32 FMA in shared memory — these problems are a few KB each and are
//     launch/latency-bound, tensor cores do not pay.
#include "common.cuh"

#include <stdlib.h>

namespace tae {

// tcgen05 path for the 256-token grid (attention_sm100.cu)
int attention_fwd_tcgen05(const bf16* qkv, bf16* out, float* lse, int B, int H, cudaStream_t stream);
int attention_bwd_tcgen05(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, const float* delta,
                          bf16* dqkv, int B, int H, cudaStream_t stream);

namespace attn {


// =============================================================================================
// Tensor-core path
// =============================================================================================
constexpr int HD = 64;
constexpr int LDS = 72;  // padded smem row (144 B): ldmatrix of 8 rows hits 32 distinct banks

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// smem byte addresses of ldmatrix row pointers (see fragment layouts of mma.m16n8k16)
// A operand: 16 rows (r0..) x 16 k (k0..) of a row-major [row][k] tile
__device__ __forceinline__ uint32_t addr_a(uint32_t base, int r0, int k0, int lane) {
  const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int col = k0 + (lane >> 4) * 8;
  return base + (uint32_t)(row * LDS + col) * 2u;
}
// B operand from smem stored [n][k] (k contiguous): two n-tiles (n0.., n0+8..) x 16 k; regs {b0,b1 | b0,b1}
__device__ __forceinline__ uint32_t addr_b(uint32_t base, int n0, int k0, int lane) {
  const int row = n0 + (lane & 7) + (lane >> 4) * 8;
  const int col = k0 + ((lane >> 3) & 1) * 8;
  return base + (uint32_t)(row * LDS + col) * 2u;
}
// B operand from smem stored [k][n] (n contiguous), loaded with .trans: 16 k (k0..) x two n-tiles (n0.., n0+8..)
__device__ __forceinline__ uint32_t addr_bt(uint32_t base, int k0, int n0, int lane) {
  const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int col = n0 + (lane >> 4) * 8;
  return base + (uint32_t)(row * LDS + col) * 2u;
}

// cooperative load of `count` matrices [N][HD] (row stride = ld_g elements in global) into padded smem
template <int N, int NTHREADS>
__device__ __forceinline__ void load_rows(bf16* s, const bf16* g, size_t ld_g, int tid) {
  for (int c = tid; c < N * 8; c += NTHREADS) {
    const int row = c >> 3, ch = c & 7;
    const uint4 v = ld_nc_v4(g + (size_t)row * ld_g + ch * 8);
    *reinterpret_cast<uint4*>(s + row * LDS + ch * 8) = v;
  }
}

// asynchronous variant (LDGSTS): the rows land in shared memory without passing through registers
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
template <int N, int NTHREADS>
__device__ __forceinline__ void load_rows_async(bf16* s, const bf16* g, size_t ld_g, int tid) {
  for (int c = tid; c < N * 8; c += NTHREADS) {
    const int row = c >> 3, ch = c & 7;
    cp_async16(smem_u32(s + row * LDS + ch * 8), g + (size_t)row * ld_g + ch * 8);
  }
}

// Persistent over (image, head) items; for N = 64 the operands are double-buffered (cp.async of item i+1 while item i
// computes), which is what the short grid needs: per-item work is a few microseconds and the one-shot version spent
// most of it waiting for its own loads.
template <int N>
__global__ void __launch_bounds__(N * 2, 1)
attn_fwd_mma(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int H, int BH,
             float scale_log2) {
  constexpr int NT = N * 2;  // threads: one warp per 16 queries
  constexpr int NBUF = N <= 64 ? 2 : 1;
  constexpr int BUF = 3 * N * LDS;  // elements of one operand buffer (Q | K | V)
  extern __shared__ uint4 smem_u4[];
  bf16* sbase = reinterpret_cast<bf16*>(smem_u4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = H * HD;
  const size_t ldq = (size_t)3 * D;
  auto issue = [&](int item, int st) {
    const int b = item / H, h = item - b * H;
    const bf16* gq = qkv + (size_t)b * N * ldq + (size_t)h * HD;
    bf16* q = sbase + st * BUF;
    load_rows_async<N, NT>(q, gq, ldq, tid);
    load_rows_async<N, NT>(q + N * LDS, gq + D, ldq, tid);
    load_rows_async<N, NT>(q + 2 * N * LDS, gq + 2 * D, ldq, tid);
  };
  int stage = 0;
  if (NBUF == 2 && (int)blockIdx.x < BH) {
    issue(blockIdx.x, 0);
    cp_async_commit();
  }
#pragma unroll 1
  for (int item = blockIdx.x; item < BH; item += gridDim.x) {
  if (NBUF == 2) {
    const int nxt = item + gridDim.x;
    if (nxt < BH) issue(nxt, stage ^ 1);
    cp_async_commit();  // (possibly empty) keeps one group per iteration
    cp_async_wait<1>();
  } else {
    issue(item, 0);
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();
  bf16* sQ = sbase + stage * BUF;
  bf16* sK = sQ + N * LDS;
  bf16* sV = sK + N * LDS;
  const int b = item / H, h = item - b * H;

  const uint32_t bQ = smem_u32(sQ), bK = smem_u32(sK), bV = smem_u32(sV);
  const int q0 = warp * 16;
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(qf[ks], addr_a(bQ, q0, ks * 16, lane));

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

#pragma unroll 1
  for (int kc = 0; kc < N / 64; ++kc) {
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kb[4];
        ldsm_x4(kb, addr_b(bK, kc * 64 + np * 16, ks * 16, lane));
        mma16816(s[2 * np], qf[ks], kb[0], kb[1]);
        mma16816(s[2 * np + 1], qf[ks], kb[2], kb[3]);
      }
    }
    // online softmax in the exp2 domain (scores pre-multiplied by scale*log2(e))
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] *= scale_log2;
      s[i][1] *= scale_log2;
      s[i][2] *= scale_log2;
      s[i][3] *= scale_log2;
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(s[i][0] - mn0), p1 = exp2f(s[i][1] - mn0);
      const float p2 = exp2f(s[i][2] - mn1), p3 = exp2f(s[i][3] - mn1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      // C fragments of n-tiles (2kk, 2kk+1) are the A fragment of k-step kk
      pf[i >> 1][(i & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i][0] *= c0;
      o[i][1] *= c0;
      o[i][2] *= c1;
      o[i][3] *= c1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        ldsm_x4_t(vb, addr_bt(bV, kc * 64 + kk * 16, dp * 16, lane));
        mma16816(o[2 * dp], pf[kk], vb[0], vb[1]);
        mma16816(o[2 * dp + 1], pf[kk], vb[2], vb[3]);
      }
    }
  }
  // finish: row sums across the quad, normalise, log-sum-exp (natural log)
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int g = lane >> 2, t = lane & 3;
  if (t == 0) {
    float* lrow = lse + ((size_t)b * H + h) * N;
    lrow[q0 + g] = (m0 + log2f(l0)) * 0.69314718055994530942f;
    lrow[q0 + g + 8] = (m1 + log2f(l1)) * 0.69314718055994530942f;
  }
  // stage this warp's 16x64 output tile through its own (now dead) Q rows, then 128-bit coalesced stores
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    *reinterpret_cast<uint32_t*>(sQ + (q0 + g) * LDS + i * 8 + 2 * t) = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + (q0 + g + 8) * LDS + i * 8 + 2 * t) = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
  }
  __syncwarp();
  bf16* go = out + (size_t)b * N * D + (size_t)h * HD;
#pragma unroll
  for (int c = lane; c < 16 * 8; c += 32) {
    const int row = q0 + (c >> 3), ch = c & 7;
    *reinterpret_cast<uint4*>(go + (size_t)row * D + ch * 8) = *reinterpret_cast<const uint4*>(sQ + row * LDS + ch * 8);
  }
  __syncthreads();  // every warp is done with this buffer before the next iteration's loads overwrite it
  if (NBUF == 2) stage ^= 1;
  }
}

// A operand (16 rows m0.. x 16 k k0..) from smem stored TRANSPOSED, [k][m] with m contiguous, loaded with .trans:
// matrices (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15) -> a0..a3
__device__ __forceinline__ uint32_t addr_at(uint32_t base, int k0, int m0, int lane) {
  const int row = k0 + (lane & 7) + (lane >> 4) * 8;
  const int col = m0 + ((lane >> 3) & 1) * 8;
  return base + (uint32_t)(row * LDS + col) * 2u;
}

// Backward for the 64-token grid (patch32: N = 64, hd = 64), persistent over (image, head) items with the operands of
// item i+1 arriving (cp.async) while item i computes.  One warp per 16 keys:
//   phase 1  S^T = K_j Q^T, dP^T = V_j dO^T for the warp's 16 keys against all 64 queries (4 steps of 16 queries);
//            P^T = exp2(S^T*c - lse2[q]), dS^T = P^T (dP^T - delta[q]); dV += P^T dO, dK += dS^T Q stay in registers;
//            the bf16 dS^T block is also written over the warp's own (dead: its V fragments are register-resident)
//            rows of the V tile, as [key][query];
//   phase 2  dQ = dS K: one warp per 16 queries, A fragments = that shared dS^T tile read transposed (ldmatrix.trans),
//            32 MMAs — instead of recomputing S and dP for every (query, key) pair a second time (96 MMAs).
// HAS_DELTA: delta = rowsum(dO * O) arrives precomputed (the proj dgrad's TAE_EPI_BF16_ROWDOT by-product), so O is
// neither read from HBM nor staged: 4 operand tiles of 9 KB, 74 KB per CTA double-buffered -> 3 CTAs per SM.  Without it
// (direct C-ABI use) O is staged as a fifth tile and delta computed from shared memory (92 KB, 2 CTAs per SM).
// No atomics, deterministic.
template <bool HAS_DELTA>
__global__ void __launch_bounds__(128, HAS_DELTA ? 3 : 2)
attn_bwd_mma64(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
               const float* __restrict__ lse, const float* __restrict__ delta_in, bf16* __restrict__ dqkv, int H, int BH,
               float scale, float scale_log2) {
  constexpr int N = 64, NT = 128;
  constexpr int NMAT = HAS_DELTA ? 4 : 5;    // Q | K | V | dO (| O)
  constexpr int BUF = NMAT * N * LDS;        // elements of one operand buffer
  extern __shared__ uint4 smem_u4[];
  bf16* sbase = reinterpret_cast<bf16*>(smem_u4);
  float* sLse = reinterpret_cast<float*>(sbase + 2 * BUF);  // lse * log2(e)
  float* sDelta = sLse + N;                                  // rowsum(dO * O)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = H * HD;
  const size_t ldq = (size_t)3 * D;
  auto issue = [&](int item, int st) {
    const int b = item / H, h = item - b * H;
    const bf16* gq = qkv + (size_t)b * N * ldq + (size_t)h * HD;
    bf16* q = sbase + st * BUF;
    load_rows_async<N, NT>(q, gq, ldq, tid);
    load_rows_async<N, NT>(q + N * LDS, gq + D, ldq, tid);
    load_rows_async<N, NT>(q + 2 * N * LDS, gq + 2 * D, ldq, tid);
    load_rows_async<N, NT>(q + 3 * N * LDS, dout + (size_t)b * N * D + (size_t)h * HD, (size_t)D, tid);
    if (!HAS_DELTA) load_rows_async<N, NT>(q + 4 * N * LDS, out + (size_t)b * N * D + (size_t)h * HD, (size_t)D, tid);
  };
  int stage = 0;
  if ((int)blockIdx.x < BH) {
    issue(blockIdx.x, 0);
    cp_async_commit();
  }
#pragma unroll 1
  for (int item = blockIdx.x; item < BH; item += gridDim.x) {
    const int nxt = item + gridDim.x;
    if (nxt < BH) issue(nxt, stage ^ 1);
    cp_async_commit();  // (possibly empty) keeps one group per iteration
    const int b = item / H, h = item - b * H;
    // per-row scalars straight from global memory while the operand copies land
    if (tid < N) {
      const size_t r = ((size_t)b * H + h) * N + tid;
      sLse[tid] = lse[r] * 1.44269504088896340736f;
      if (HAS_DELTA) sDelta[tid] = delta_in[r];
    }
    cp_async_wait<1>();
    __syncthreads();
    bf16* sQ = sbase + stage * BUF;
    bf16* sK = sQ + N * LDS;
    bf16* sV = sK + N * LDS;
    bf16* sdO = sV + N * LDS;
    if (!HAS_DELTA) {
      // delta = rowsum(dO * O) out of shared memory: 8 consecutive lanes own the 8 16-byte chunks of one row
      const bf16* sO = sdO + N * LDS;
      for (int c = tid; c < N * 8; c += NT) {
        const int row = c >> 3, ch = c & 7;
        const uint4 dv = *reinterpret_cast<const uint4*>(sdO + row * LDS + ch * 8);
        const uint4 ov = *reinterpret_cast<const uint4*>(sO + row * LDS + ch * 8);
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, ow[4] = {ov.x, ov.y, ov.z, ov.w};
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 d2 = unpack_bf16x2(dw[j]), o2 = unpack_bf16x2(ow[j]);
          acc += d2.x * o2.x + d2.y * o2.y;
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (ch == 0) sDelta[row] = acc;
      }
      __syncthreads();
    }

    const uint32_t bQ = smem_u32(sQ), bK = smem_u32(sK), bV = smem_u32(sV), bdO = smem_u32(sdO);
    const int g = lane >> 2, t = lane & 3;
    bf16* gdq = dqkv + (size_t)b * N * ldq + (size_t)h * HD;

    // ---------------- phase 1: this warp owns keys j0..j0+15 -> dK, dV, and the dS^T rows of those keys ----------------
    {
      const int j0 = warp * 16;
      uint32_t kf[4][4], vf[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        ldsm_x4(kf[ks], addr_a(bK, j0, ks * 16, lane));
        ldsm_x4(vf[ks], addr_a(bV, j0, ks * 16, lane));
      }
      __syncwarp();  // every lane holds its V fragments: rows j0..j0+15 of the V tile are dead from here on
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
      }
#pragma unroll 1
      for (int i0 = 0; i0 < N; i0 += 16) {
        // S^T[key, query] = K_j Q_i^T ; dP^T[key, query] = V_j dO_i^T
        float st[2][4], dpt[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
          dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t qb[4], ob[4];
          ldsm_x4(qb, addr_b(bQ, i0, ks * 16, lane));
          ldsm_x4(ob, addr_b(bdO, i0, ks * 16, lane));
          mma16816(st[0], kf[ks], qb[0], qb[1]);
          mma16816(st[1], kf[ks], qb[2], qb[3]);
          mma16816(dpt[0], vf[ks], ob[0], ob[1]);
          mma16816(dpt[1], vf[ks], ob[2], ob[3]);
        }
        // P^T = exp2(S^T*scale*log2e - lse2[query]); dS^T = P^T * (dP^T - delta[query])   (queries index columns)
        uint32_t pa[4], dsa[4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int qc = i0 + nt * 8 + 2 * t;
          const float2 l01 = *reinterpret_cast<const float2*>(sLse + qc), d01 = *reinterpret_cast<const float2*>(sDelta + qc);
          const float p0 = exp2f(fmaf(st[nt][0], scale_log2, -l01.x)), p1 = exp2f(fmaf(st[nt][1], scale_log2, -l01.y));
          const float p2 = exp2f(fmaf(st[nt][2], scale_log2, -l01.x)), p3 = exp2f(fmaf(st[nt][3], scale_log2, -l01.y));
          pa[nt * 2 + 0] = pack_bf16x2(p0, p1);
          pa[nt * 2 + 1] = pack_bf16x2(p2, p3);
          dsa[nt * 2 + 0] = pack_bf16x2(p0 * (dpt[nt][0] - d01.x), p1 * (dpt[nt][1] - d01.y));
          dsa[nt * 2 + 1] = pack_bf16x2(p2 * (dpt[nt][2] - d01.x), p3 * (dpt[nt][3] - d01.y));
          // dS^T block -> this warp's rows of the V tile, [key][query]: rows j0+g / j0+g+8, columns qc, qc+1
          *reinterpret_cast<uint32_t*>(sV + (j0 + g) * LDS + qc) = dsa[nt * 2 + 0];
          *reinterpret_cast<uint32_t*>(sV + (j0 + g + 8) * LDS + qc) = dsa[nt * 2 + 1];
        }
        // dV += P^T dO_i ; dK += dS^T Q_i     (k = 16 queries; B operands [query][hd] loaded transposed)
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t ob[4], qb[4];
          ldsm_x4_t(ob, addr_bt(bdO, i0, dp * 16, lane));
          mma16816(dv[2 * dp], pa, ob[0], ob[1]);
          mma16816(dv[2 * dp + 1], pa, ob[2], ob[3]);
          ldsm_x4_t(qb, addr_bt(bQ, i0, dp * 16, lane));
          mma16816(dk[2 * dp], dsa, qb[0], qb[1]);
          mma16816(dk[2 * dp + 1], dsa, qb[2], qb[3]);
        }
      }
      bf16* gdk = gdq + D;
      bf16* gdv = gdq + 2 * D;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = i * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(gdk + (size_t)(j0 + g) * ldq + col) = pack_bf16x2(dk[i][0] * scale, dk[i][1] * scale);
        *reinterpret_cast<uint32_t*>(gdk + (size_t)(j0 + g + 8) * ldq + col) = pack_bf16x2(dk[i][2] * scale, dk[i][3] * scale);
        *reinterpret_cast<uint32_t*>(gdv + (size_t)(j0 + g) * ldq + col) = pack_bf16x2(dv[i][0], dv[i][1]);
        *reinterpret_cast<uint32_t*>(gdv + (size_t)(j0 + g + 8) * ldq + col) = pack_bf16x2(dv[i][2], dv[i][3]);
      }
    }
    __syncthreads();  // the whole dS^T tile [64 keys][64 queries] is in shared memory

    // ---------------- phase 2: this warp owns queries i0..i0+15 -> dQ = dS K ----------------
    {
      const int i0 = warp * 16;
      float dq[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step
        uint32_t dsf[4];
        ldsm_x4_t(dsf, addr_at(bV, ks * 16, i0, lane));
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t kb[4];
          ldsm_x4_t(kb, addr_bt(bK, ks * 16, dp * 16, lane));
          mma16816(dq[2 * dp], dsf, kb[0], kb[1]);
          mma16816(dq[2 * dp + 1], dsf, kb[2], kb[3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = i * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(gdq + (size_t)(i0 + g) * ldq + col) = pack_bf16x2(dq[i][0] * scale, dq[i][1] * scale);
        *reinterpret_cast<uint32_t*>(gdq + (size_t)(i0 + g + 8) * ldq + col) = pack_bf16x2(dq[i][2] * scale, dq[i][3] * scale);
      }
    }
    __syncthreads();  // every warp is done with this buffer (and lse/delta) before the next iteration overwrites them
    stage ^= 1;
  }
}

// =============================================================================================
// Tiny grids (N <= 16 tokens: patch64 / patch128 models, hd = 80): ONE WARP per (image, head), tensor-core math
// (mma.sync m16n8k16) on a 16x16 score tile.  Rows / keys >= N are zero-padded and masked.  The fp32 shared-memory
// version of this case was bound by shared-memory bandwidth (110 us forward at B=256, H=32, N=16 against a 13 us HBM
// floor).  The backward needs no saved output: delta = rowsum(dO * O) = rowsum(P * dP).
// =============================================================================================
__device__ __forceinline__ uint32_t movmatrix_t(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

template <int HDT>
struct Small {
  static constexpr int LD = HDT + 8;      // padded row (hd = 80: 176 B -> 8 ldmatrix rows on distinct banks)
  static constexpr int KS = HDT / 16;     // k-steps over the head dimension
  static constexpr int NT8 = HDT / 8;     // 8-wide n-tiles over the head dimension
  static constexpr int TILE = 16 * LD;    // elements of one [16 x HDT] tile
  __device__ static uint32_t a_addr(uint32_t base, int k0, int lane) {  // A: rows 0..15, k0..k0+15 of [row][k]
    return base + (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * LD + k0 + (lane >> 4) * 8) * 2u;
  }
  __device__ static uint32_t b_addr(uint32_t base, int k0, int lane) {  // B from [n][k]: n 0..15 (two n-tiles), k0..
    return base + (uint32_t)(((lane & 7) + (lane >> 4) * 8) * LD + k0 + ((lane >> 3) & 1) * 8) * 2u;
  }
  __device__ static uint32_t bt_addr(uint32_t base, int n0, int lane) {  // B from [k][n] (.trans): k 0..15, n0..n0+15
    return base + (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * LD + n0 + (lane >> 4) * 8) * 2u;
  }
  // P heads x [N x HDT] rows of global (row stride ld_g, heads HDT columns apart) -> padded smem tile: tile row r holds
  // token r % N of head r / N; rows >= P * N zeroed
  __device__ static void load_tile(bf16* s, const bf16* g, size_t ld_g, int N, int PN, int lane) {
    constexpr int CPR = HDT / 8;
    for (int c = lane; c < 16 * CPR; c += 32) {
      const int row = c / CPR, ch = c - row * CPR;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (row < PN) {
        const int hh = row / N, tok = row - hh * N;
        v = ld_nc_v4(g + (size_t)tok * ld_g + hh * HDT + ch * 8);
      }
      *reinterpret_cast<uint4*>(s + row * LD + ch * 8) = v;
    }
  }
};

// S (exp2 domain, masked) of one 16x16 tile: s[nt][0..3] = rows (g, g+8), keys nt*8 + 2t + {0,1}.  The tile packs
// PN / N heads of N tokens each: a key counts for a query only inside the same head (block-diagonal mask).
template <int HDT>
__device__ __forceinline__ void small_scores(float (&s)[2][4], uint32_t bQ, uint32_t bK, int N, int PN, float scale_log2,
                                             int lane) {
  using C = Small<HDT>;
#pragma unroll
  for (int i = 0; i < 2; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < C::KS; ++ks) {
    uint32_t qa[4], kb[4];
    ldsm_x4(qa, C::a_addr(bQ, ks * 16, lane));
    ldsm_x4(kb, C::b_addr(bK, ks * 16, lane));
    mma16816(s[0], qa, kb[0], kb[1]);
    mma16816(s[1], qa, kb[2], kb[3]);
  }
  const int t = lane & 3, g = lane >> 2;
  const int blk0 = g / N, blk1 = (g + 8) / N;  // head (block) of this lane's two query rows
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int j0 = nt * 8 + 2 * t, j1 = j0 + 1;
    const int kb0 = j0 < PN ? j0 / N : -1, kb1 = j1 < PN ? j1 / N : -1;  // -1: padding key
    s[nt][0] = kb0 == blk0 ? s[nt][0] * scale_log2 : -INFINITY;
    s[nt][2] = kb0 == blk1 ? s[nt][2] * scale_log2 : -INFINITY;
    s[nt][1] = kb1 == blk0 ? s[nt][1] * scale_log2 : -INFINITY;
    s[nt][3] = kb1 == blk1 ? s[nt][3] * scale_log2 : -INFINITY;
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

template <int HDT>
__global__ void __launch_bounds__(128)
attn_fwd_small(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int items, int N, int H,
               int P, float scale_log2) {
  using C = Small<HDT>;
  extern __shared__ uint4 smem_u4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * 4 + warp;
  if (wg >= items) return;  // no CTA-wide barriers below
  // one warp = P consecutive heads of one image (P = 16 / N when that divides the heads: 4 heads of the 4-token grid),
  // packed as a block-diagonal 16x16 problem
  const int gpi = H / P, b = wg / gpi, h = (wg - b * gpi) * P, PN = P * N;
  const int D = H * HDT;
  const size_t ldq = (size_t)3 * D;
  bf16* sQ = reinterpret_cast<bf16*>(smem_u4) + (size_t)warp * 3 * C::TILE;
  bf16* sK = sQ + C::TILE;
  bf16* sV = sK + C::TILE;
  const bf16* gq = qkv + (size_t)b * N * ldq + (size_t)h * HDT;
  C::load_tile(sQ, gq, ldq, N, PN, lane);
  C::load_tile(sK, gq + D, ldq, N, PN, lane);
  C::load_tile(sV, gq + 2 * D, ldq, N, PN, lane);
  __syncwarp();
  const uint32_t bQ = smem_u32(sQ), bK = smem_u32(sK), bV = smem_u32(sV);
  float s[2][4];
  small_scores<HDT>(s, bQ, bK, N, PN, scale_log2, lane);
  const float m0 = quad_max(fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1])));
  const float m1 = quad_max(fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3])));
  float p[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    p[nt][0] = exp2f(s[nt][0] - m0);
    p[nt][1] = exp2f(s[nt][1] - m0);
    p[nt][2] = exp2f(s[nt][2] - m1);
    p[nt][3] = exp2f(s[nt][3] - m1);
  }
  const float l0 = quad_sum(p[0][0] + p[0][1] + p[1][0] + p[1][1]);
  const float l1 = quad_sum(p[0][2] + p[0][3] + p[1][2] + p[1][3]);
  // C fragments of the two key n-tiles are the A fragment of the single 16-key k-step of P V
  const uint32_t pa[4] = {pack_bf16x2(p[0][0], p[0][1]), pack_bf16x2(p[0][2], p[0][3]), pack_bf16x2(p[1][0], p[1][1]),
                          pack_bf16x2(p[1][2], p[1][3])};
  const int g = lane >> 2, t = lane & 3;
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  // tile rows g and g + 8 -> (head offset, token)
  const int hh0 = g / N, tok0 = g - hh0 * N, hh1 = (g + 8) / N, tok1 = g + 8 - hh1 * N;
  if (t == 0) {
    float* lrow = lse + ((size_t)b * H + h) * N;  // [B, H, N]: head h + hh at offset hh * N
    if (g < PN) lrow[hh0 * N + tok0] = (m0 + log2f(l0)) * 0.69314718055994530942f;
    if (g + 8 < PN) lrow[hh1 * N + tok1] = (m1 + log2f(l1)) * 0.69314718055994530942f;
  }
  bf16* go = out + (size_t)b * N * D + (size_t)h * HDT;
  bf16* go0 = go + (size_t)tok0 * D + hh0 * HDT;
  bf16* go1 = go + (size_t)tok1 * D + hh1 * HDT;
#pragma unroll
  for (int dp = 0; dp < C::NT8 / 2; ++dp) {
    uint32_t vb[4];
    ldsm_x4_t(vb, C::bt_addr(bV, dp * 16, lane));
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    mma16816(o0, pa, vb[0], vb[1]);
    mma16816(o1, pa, vb[2], vb[3]);
    const int col = dp * 16 + 2 * t;
    if (g < PN) {
      *reinterpret_cast<uint32_t*>(go0 + col) = pack_bf16x2(o0[0] * inv0, o0[1] * inv0);
      *reinterpret_cast<uint32_t*>(go0 + col + 8) = pack_bf16x2(o1[0] * inv0, o1[1] * inv0);
    }
    if (g + 8 < PN) {
      *reinterpret_cast<uint32_t*>(go1 + col) = pack_bf16x2(o0[2] * inv1, o0[3] * inv1);
      *reinterpret_cast<uint32_t*>(go1 + col + 8) = pack_bf16x2(o1[2] * inv1, o1[3] * inv1);
    }
  }
}

template <int HDT>
__global__ void __launch_bounds__(128)
attn_bwd_small(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
               bf16* __restrict__ dqkv, int items, int N, int H, int P, float scale, float scale_log2) {
  using C = Small<HDT>;
  extern __shared__ uint4 smem_u4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * 4 + warp;
  if (wg >= items) return;
  const int gpi = H / P, b = wg / gpi, h = (wg - b * gpi) * P, PN = P * N;  // P heads per warp, see attn_fwd_small
  const int D = H * HDT;
  const size_t ldq = (size_t)3 * D;
  bf16* sQ = reinterpret_cast<bf16*>(smem_u4) + (size_t)warp * 4 * C::TILE;
  bf16* sK = sQ + C::TILE;
  bf16* sV = sK + C::TILE;
  bf16* sdO = sV + C::TILE;
  const bf16* gq = qkv + (size_t)b * N * ldq + (size_t)h * HDT;
  C::load_tile(sQ, gq, ldq, N, PN, lane);
  C::load_tile(sK, gq + D, ldq, N, PN, lane);
  C::load_tile(sV, gq + 2 * D, ldq, N, PN, lane);
  C::load_tile(sdO, dout + (size_t)b * N * D + (size_t)h * HDT, (size_t)D, N, PN, lane);
  __syncwarp();
  const uint32_t bQ = smem_u32(sQ), bK = smem_u32(sK), bV = smem_u32(sV), bdO = smem_u32(sdO);
  const int g = lane >> 2, t = lane & 3;
  // P = exp2(S * c - lse2[row])   (masked keys give exp2(-inf) = 0; padded rows are discarded at the stores)
  float s[2][4];
  small_scores<HDT>(s, bQ, bK, N, PN, scale_log2, lane);
  const int hh0 = g / N, tok0 = g - hh0 * N, hh1 = (g + 8) / N, tok1 = g + 8 - hh1 * N;  // tile rows g, g + 8
  const float* lrow = lse + ((size_t)b * H + h) * N;
  const float ls0 = g < PN ? lrow[hh0 * N + tok0] * 1.44269504088896340736f : 0.f;
  const float ls1 = g + 8 < PN ? lrow[hh1 * N + tok1] * 1.44269504088896340736f : 0.f;
  float p[2][4], dp[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    p[nt][0] = exp2f(s[nt][0] - ls0);
    p[nt][1] = exp2f(s[nt][1] - ls0);
    p[nt][2] = exp2f(s[nt][2] - ls1);
    p[nt][3] = exp2f(s[nt][3] - ls1);
    dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
  }
  // dP = dO V^T
#pragma unroll
  for (int ks = 0; ks < C::KS; ++ks) {
    uint32_t da[4], vb[4];
    ldsm_x4(da, C::a_addr(bdO, ks * 16, lane));
    ldsm_x4(vb, C::b_addr(bV, ks * 16, lane));
    mma16816(dp[0], da, vb[0], vb[1]);
    mma16816(dp[1], da, vb[2], vb[3]);
  }
  // delta = rowsum(P * dP)  (== rowsum(dO * O));  dS = P (dP - delta) * scale
  const float dl0 = quad_sum(p[0][0] * dp[0][0] + p[0][1] * dp[0][1] + p[1][0] * dp[1][0] + p[1][1] * dp[1][1]);
  const float dl1 = quad_sum(p[0][2] * dp[0][2] + p[0][3] * dp[0][3] + p[1][2] * dp[1][2] + p[1][3] * dp[1][3]);
  float ds[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    ds[nt][0] = p[nt][0] * (dp[nt][0] - dl0) * scale;
    ds[nt][1] = p[nt][1] * (dp[nt][1] - dl0) * scale;
    ds[nt][2] = p[nt][2] * (dp[nt][2] - dl1) * scale;
    ds[nt][3] = p[nt][3] * (dp[nt][3] - dl1) * scale;
  }
  // A fragments: dS [q x key] for dQ; dS^T and P^T [key x q] (8x8 blocks transposed with movmatrix) for dK and dV
  const uint32_t dsa[4] = {pack_bf16x2(ds[0][0], ds[0][1]), pack_bf16x2(ds[0][2], ds[0][3]), pack_bf16x2(ds[1][0], ds[1][1]),
                           pack_bf16x2(ds[1][2], ds[1][3])};
  const uint32_t pa[4] = {pack_bf16x2(p[0][0], p[0][1]), pack_bf16x2(p[0][2], p[0][3]), pack_bf16x2(p[1][0], p[1][1]),
                          pack_bf16x2(p[1][2], p[1][3])};
  const uint32_t dst_[4] = {movmatrix_t(dsa[0]), movmatrix_t(dsa[2]), movmatrix_t(dsa[1]), movmatrix_t(dsa[3])};
  const uint32_t pt_[4] = {movmatrix_t(pa[0]), movmatrix_t(pa[2]), movmatrix_t(pa[1]), movmatrix_t(pa[3])};
  bf16* gd = dqkv + (size_t)b * N * ldq + (size_t)h * HDT;
#pragma unroll
  for (int dpi = 0; dpi < C::NT8 / 2; ++dpi) {
    uint32_t kb[4], qb[4], ob[4];
    ldsm_x4_t(kb, C::bt_addr(bK, dpi * 16, lane));    // K  [key][d]
    ldsm_x4_t(qb, C::bt_addr(bQ, dpi * 16, lane));    // Q  [q][d]
    ldsm_x4_t(ob, C::bt_addr(bdO, dpi * 16, lane));   // dO [q][d]
    float q0[4] = {0.f, 0.f, 0.f, 0.f}, q1[4] = {0.f, 0.f, 0.f, 0.f};
    float k0[4] = {0.f, 0.f, 0.f, 0.f}, k1[4] = {0.f, 0.f, 0.f, 0.f};
    float v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f};
    mma16816(q0, dsa, kb[0], kb[1]);   // dQ = dS K
    mma16816(q1, dsa, kb[2], kb[3]);
    mma16816(k0, dst_, qb[0], qb[1]);  // dK = dS^T Q
    mma16816(k1, dst_, qb[2], qb[3]);
    mma16816(v0, pt_, ob[0], ob[1]);   // dV = P^T dO
    mma16816(v1, pt_, ob[2], ob[3]);
    const int col = dpi * 16 + 2 * t;
    if (g < PN) {
      bf16* r = gd + (size_t)tok0 * ldq + hh0 * HDT + col;
      *reinterpret_cast<uint32_t*>(r) = pack_bf16x2(q0[0], q0[1]);
      *reinterpret_cast<uint32_t*>(r + 8) = pack_bf16x2(q1[0], q1[1]);
      *reinterpret_cast<uint32_t*>(r + D) = pack_bf16x2(k0[0], k0[1]);
      *reinterpret_cast<uint32_t*>(r + D + 8) = pack_bf16x2(k1[0], k1[1]);
      *reinterpret_cast<uint32_t*>(r + 2 * D) = pack_bf16x2(v0[0], v0[1]);
      *reinterpret_cast<uint32_t*>(r + 2 * D + 8) = pack_bf16x2(v1[0], v1[1]);
    }
    if (g + 8 < PN) {
      bf16* r = gd + (size_t)tok1 * ldq + hh1 * HDT + col;
      *reinterpret_cast<uint32_t*>(r) = pack_bf16x2(q0[2], q0[3]);
      *reinterpret_cast<uint32_t*>(r + 8) = pack_bf16x2(q1[2], q1[3]);
      *reinterpret_cast<uint32_t*>(r + D) = pack_bf16x2(k0[2], k0[3]);
      *reinterpret_cast<uint32_t*>(r + D + 8) = pack_bf16x2(k1[2], k1[3]);
      *reinterpret_cast<uint32_t*>(r + 2 * D) = pack_bf16x2(v0[2], v0[3]);
      *reinterpret_cast<uint32_t*>(r + 2 * D + 8) = pack_bf16x2(v1[2], v1[3]);
    }
  }
}

template <typename K>
static int set_smem(K kernel, int bytes);

template <int HDT>
static int launch_small(bool bwd, const bf16* qkv, const bf16* dout, bf16* out_or_dqkv, const float* lse_in, float* lse_out,
                        int B, int N, int H, float scale, cudaStream_t stream) {
  // heads per warp: the 16-row tile holds 16 / N heads of one image when that count divides the heads (4-token grid of
  // patch128: 4 heads per warp — a quarter of the warps, loads and stores of the one-head-per-warp layout)
  int P = 1;
  if (N < 16 && 16 % N == 0 && H % (16 / N) == 0) P = 16 / N;
  const int items = B * (H / P);
  const int smem = 4 * (bwd ? 4 : 3) * Small<HDT>::TILE * 2;  // 4 warps per CTA
  const float sl2 = scale * 1.44269504088896340736f;
  if (!bwd) {
    const int rc = set_smem(attn_fwd_small<HDT>, smem);
    if (rc) return rc;
    attn_fwd_small<HDT><<<(items + 3) / 4, 128, smem, stream>>>(qkv, out_or_dqkv, lse_out, items, N, H, P, sl2);
  } else {
    const int rc = set_smem(attn_bwd_small<HDT>, smem);
    if (rc) return rc;
    attn_bwd_small<HDT><<<(items + 3) / 4, 128, smem, stream>>>(qkv, dout, lse_in, out_or_dqkv, items, N, H, P, scale, sl2);
  }
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
// N <= 16 and a head dimension the tensor-core kernels are instantiated for
static bool small_ok(int N, int hd) { return N <= 16 && (hd == 32 || hd == 64 || hd == 80); }
static int dispatch_small(bool bwd, const bf16* qkv, const bf16* dout, bf16* o, const float* lse_in, float* lse_out, int B,
                          int N, int H, int hd, float scale, cudaStream_t stream) {
  if (hd == 32) return launch_small<32>(bwd, qkv, dout, o, lse_in, lse_out, B, N, H, scale, stream);
  if (hd == 64) return launch_small<64>(bwd, qkv, dout, o, lse_in, lse_out, B, N, H, scale, stream);
  return launch_small<80>(bwd, qkv, dout, o, lse_in, lse_out, B, N, H, scale, stream);
}

// =============================================================================================
// Generic path: one warp per (image, head), fp32 math in shared memory
// =============================================================================================
__device__ __forceinline__ void load_head_f32(float* s, const bf16* g, size_t ld_g, int N, int hd, int ldp, int lane) {
  const int cpr = hd >> 3;  // 16-byte chunks per row
  for (int c = lane; c < N * cpr; c += 32) {
    const int row = c / cpr, ch = c - row * cpr;
    const uint4 v = ld_nc_v4(g + (size_t)row * ld_g + ch * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float* dst = s + row * ldp + ch * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      dst[2 * j] = f.x;
      dst[2 * j + 1] = f.y;
    }
  }
}

// per-warp smem floats: fwd 3*N*ldp + N*(N+1); bwd 4*N*ldp + 2*N*(N+1) + N, with ldp = hd + 4: rows are 16-byte
// aligned so every dot product runs on 128-bit shared loads (the scalar version was LDS-bound: 2 loads per FMA), and
// the 4-float skew keeps the 8 rows a quarter-warp touches on distinct banks (hd = 80: starts 0,20,8,28,16,4,24,12).
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ void axpy4(float4& acc, float s, const float4 v) {
  acc.x = fmaf(s, v.x, acc.x);
  acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z);
  acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ void store_bf16x4(bf16* dst, const float4 v) {
  *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

__global__ void attn_fwd_simt(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int BH,
                              int N, int H, int hd, float scale, int per_warp_floats) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + warp;
  if (wg >= BH) return;  // no CTA-wide barriers below
  const int b = wg / H, h = wg % H;
  const int D = H * hd, ldp = hd + 4, ldn = N + 1, h4 = hd >> 2;
  const size_t ldq = (size_t)3 * D;
  float* sq = smem_f + (size_t)warp * per_warp_floats;
  float* sk = sq + N * ldp;
  float* sv = sk + N * ldp;
  float* sp = sv + N * ldp;
  const bf16* gq = qkv + (size_t)b * N * ldq + (size_t)h * hd;
  load_head_f32(sq, gq, ldq, N, hd, ldp, lane);
  load_head_f32(sk, gq + D, ldq, N, hd, ldp, lane);
  load_head_f32(sv, gq + 2 * D, ldq, N, hd, ldp, lane);
  __syncwarp();
  for (int e = lane; e < N * N; e += 32) {
    const int i = e / N, j = e - i * N;
    const float4* a = reinterpret_cast<const float4*>(sq + i * ldp);
    const float4* c = reinterpret_cast<const float4*>(sk + j * ldp);
    float acc0 = 0.f, acc1 = 0.f;
    int d = 0;
    for (; d + 1 < h4; d += 2) {
      acc0 = dot4(a[d], c[d], acc0);
      acc1 = dot4(a[d + 1], c[d + 1], acc1);
    }
    if (d < h4) acc0 = dot4(a[d], c[d], acc0);
    sp[i * ldn + j] = (acc0 + acc1) * scale;
  }
  __syncwarp();
  for (int i = lane; i < N; i += 32) {
    float* row = sp + i * ldn;
    float m = -INFINITY;
    for (int j = 0; j < N; ++j) m = fmaxf(m, row[j]);
    float l = 0.f;
    for (int j = 0; j < N; ++j) {
      const float p = __expf(row[j] - m);
      row[j] = p;
      l += p;
    }
    const float inv = 1.f / l;
    for (int j = 0; j < N; ++j) row[j] *= inv;
    lse[((size_t)b * H + h) * N + i] = m + __logf(l);
  }
  __syncwarp();
  bf16* go = out + (size_t)b * N * D + (size_t)h * hd;
  for (int e = lane; e < N * h4; e += 32) {  // 4 output columns per lane: one scalar + one 128-bit load per 4 FMAs
    const int i = e / h4, d4 = e - i * h4;
    const float* p = sp + i * ldn;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < N; ++j) axpy4(acc, p[j], reinterpret_cast<const float4*>(sv + j * ldp)[d4]);
    store_bf16x4(go + (size_t)i * D + d4 * 4, acc);
  }
}

__global__ void attn_bwd_simt(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                              bf16* __restrict__ dqkv, int BH, int N, int H, int hd, float scale, int per_warp_floats) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (blockDim.x >> 5) + warp;
  if (wg >= BH) return;
  const int b = wg / H, h = wg % H;
  const int D = H * hd, ldp = hd + 4, ldn = N + 1, h4 = hd >> 2;
  const size_t ldq = (size_t)3 * D;
  float* sq = smem_f + (size_t)warp * per_warp_floats;
  float* sk = sq + N * ldp;
  float* sv = sk + N * ldp;
  float* sdo = sv + N * ldp;
  float* sp = sdo + N * ldp;   // P
  float* sds = sp + N * ldn;   // dP, then dS
  float* sdelta = sds + N * ldn;
  const bf16* gq = qkv + (size_t)b * N * ldq + (size_t)h * hd;
  load_head_f32(sq, gq, ldq, N, hd, ldp, lane);
  load_head_f32(sk, gq + D, ldq, N, hd, ldp, lane);
  load_head_f32(sv, gq + 2 * D, ldq, N, hd, ldp, lane);
  load_head_f32(sdo, dout + (size_t)b * N * D + (size_t)h * hd, (size_t)D, N, hd, ldp, lane);
  __syncwarp();
  const float* lrow = lse + ((size_t)b * H + h) * N;
  for (int e = lane; e < N * N; e += 32) {
    const int i = e / N, j = e - i * N;
    const float4* qa = reinterpret_cast<const float4*>(sq + i * ldp);
    const float4* ka = reinterpret_cast<const float4*>(sk + j * ldp);
    const float4* da = reinterpret_cast<const float4*>(sdo + i * ldp);
    const float4* va = reinterpret_cast<const float4*>(sv + j * ldp);
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < h4; ++d) {
      s = dot4(qa[d], ka[d], s);
      dp = dot4(da[d], va[d], dp);
    }
    sp[i * ldn + j] = __expf(s * scale - lrow[i]);
    sds[i * ldn + j] = dp;
  }
  __syncwarp();
  for (int i = lane; i < N; i += 32) {
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(sp[i * ldn + j], sds[i * ldn + j], acc);
    sdelta[i] = acc;
  }
  __syncwarp();
  for (int e = lane; e < N * N; e += 32) {
    const int i = e / N, j = e - i * N;
    sds[i * ldn + j] = sp[i * ldn + j] * (sds[i * ldn + j] - sdelta[i]) * scale;
  }
  __syncwarp();
  bf16* gdq = dqkv + (size_t)b * N * ldq + (size_t)h * hd;
  for (int e = lane; e < N * h4; e += 32) {
    const int r = e / h4, d4 = e - r * h4;
    float4 aq = make_float4(0.f, 0.f, 0.f, 0.f), ak = aq, av = aq;
    for (int j = 0; j < N; ++j) {
      axpy4(aq, sds[r * ldn + j], reinterpret_cast<const float4*>(sk + j * ldp)[d4]);   // dQ[r] = sum_j dS[r,j] K[j]
      axpy4(ak, sds[j * ldn + r], reinterpret_cast<const float4*>(sq + j * ldp)[d4]);   // dK[r] = sum_i dS[i,r] Q[i]
      axpy4(av, sp[j * ldn + r], reinterpret_cast<const float4*>(sdo + j * ldp)[d4]);   // dV[r] = sum_i P[i,r] dO[i]
    }
    bf16* dst = gdq + (size_t)r * ldq + d4 * 4;
    store_bf16x4(dst, aq);
    store_bf16x4(dst + D, ak);
    store_bf16x4(dst + 2 * D, av);
  }
}

constexpr int SIMT_MAX_SMEM = 200 * 1024;

static int simt_config(int N, int hd, bool bwd, int* wpc, int* per_warp_floats) {
  const int ldp = hd + 4, ldn = N + 1;
  const int pw = ((bwd ? (4 * N * ldp + 2 * N * ldn + N) : (3 * N * ldp + N * ldn)) + 3) & ~3;  // 16-byte multiple
  const int bytes = pw * 4;
  if (bytes > SIMT_MAX_SMEM) return -1;
  int w = SIMT_MAX_SMEM / bytes;
  if (w > 4) w = 4;
  // keep several CTAs per SM when the problem is tiny
  while (w > 1 && w * bytes > 96 * 1024) --w;
  *wpc = w;
  *per_warp_floats = pw;
  return 0;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d) failed: %s", bytes, cudaGetErrorString(e));
      return TAE_ERR_CUDA;
    }
  }
  return TAE_OK;
}

}  // namespace attn
}  // namespace tae

namespace tae {
namespace attn {
// N = 64, hd = 64: delta != NULL selects the 4-tile kernel (3 CTAs per SM), otherwise O is staged and delta computed
static int launch_bwd64(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, const float* delta, bf16* dqkv,
                        int B, int H, float scale, cudaStream_t stream) {
  constexpr int N = 64;
  const float sl2 = scale * 1.44269504088896340736f;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  if (delta != nullptr) {
    const int smem = 2 * 4 * N * LDS * 2 + 2 * N * 4;  // 74 240 B: three CTAs per SM
    int rc = set_smem(attn_bwd_mma64<true>, smem);
    if (rc) return rc;
    const int grid = B * H < sms * 3 ? B * H : sms * 3;
    attn_bwd_mma64<true><<<grid, 128, smem, stream>>>(qkv, nullptr, dout, lse, delta, dqkv, H, B * H, scale, sl2);
  } else {
    const int smem = 2 * 5 * N * LDS * 2 + 2 * N * 4;  // 92 672 B: two CTAs per SM
    int rc = set_smem(attn_bwd_mma64<false>, smem);
    if (rc) return rc;
    const int grid = B * H < sms * 2 ? B * H : sms * 2;
    attn_bwd_mma64<false><<<grid, 128, smem, stream>>>(qkv, out, dout, lse, nullptr, dqkv, H, B * H, scale, sl2);
  }
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}
}  // namespace attn
}  // namespace tae

extern "C" int tae_attention_fwd(const tae_bf16* qkv_, tae_bf16* out_, float* lse, int32_t B, int32_t N, int32_t H,
                                 int32_t hd, void* stream_) {
  using namespace tae;
  using namespace tae::attn;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bf16* qkv = reinterpret_cast<const bf16*>(qkv_);
  bf16* out = reinterpret_cast<bf16*>(out_);
  TAE_CHECK_SHAPE(B > 0 && N > 0 && H > 0 && hd > 0, "tae_attention_fwd: non-positive dims");
  TAE_CHECK_SHAPE(hd % 8 == 0 && hd <= 128, "tae_attention_fwd: hd=%d unsupported (need hd %% 8 == 0, hd <= 128)", hd);
  const float scale = 1.0f / sqrtf((float)hd);
  if (hd == HD && N == 256) return attention_fwd_tcgen05(qkv, out, lse, B, H, stream);
  if (hd == HD && N == 64) {
    const float sl2 = scale * 1.44269504088896340736f;
    const int smem2 = 2 * 3 * N * LDS * 2;  // double-buffered operands
    int rc = set_smem(attn_fwd_mma<64>, smem2);
    if (rc) return rc;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    const int grid = B * H < sms * 3 ? B * H : sms * 3;  // 3 resident CTAs per SM (140 registers, 54 KB each)
    attn_fwd_mma<64><<<grid, 128, smem2, stream>>>(qkv, out, lse, H, B * H, sl2);
    TAE_CHECK_LAUNCH();
    return TAE_OK;
  }
  if (small_ok(N, hd)) return dispatch_small(false, qkv, nullptr, out, nullptr, lse, B, N, H, hd, scale, stream);
  int wpc, pw;
  TAE_CHECK_SHAPE(simt_config(N, hd, false, &wpc, &pw) == 0, "tae_attention_fwd: N=%d hd=%d does not fit shared memory", N, hd);
  const int smem = wpc * pw * 4;
  int rc = set_smem(attn_fwd_simt, smem);
  if (rc) return rc;
  const int BH = B * H;
  attn_fwd_simt<<<(BH + wpc - 1) / wpc, wpc * 32, smem, stream>>>(qkv, out, lse, BH, N, H, hd, scale, pw);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_attention_bwd(const tae_bf16* qkv_, const tae_bf16* out_, const tae_bf16* dout_, const float* lse,
                                 tae_bf16* dqkv_, int32_t B, int32_t N, int32_t H, int32_t hd, void* stream_) {
  using namespace tae;
  using namespace tae::attn;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bf16* qkv = reinterpret_cast<const bf16*>(qkv_);
  const bf16* out = reinterpret_cast<const bf16*>(out_);
  const bf16* dout = reinterpret_cast<const bf16*>(dout_);
  bf16* dqkv = reinterpret_cast<bf16*>(dqkv_);
  TAE_CHECK_SHAPE(B > 0 && N > 0 && H > 0 && hd > 0, "tae_attention_bwd: non-positive dims");
  TAE_CHECK_SHAPE(hd % 8 == 0 && hd <= 128, "tae_attention_bwd: hd=%d unsupported", hd);
  const float scale = 1.0f / sqrtf((float)hd);
  if (hd == HD && N == 256) return attention_bwd_tcgen05(qkv, out, dout, lse, nullptr, dqkv, B, H, stream);
  if (hd == HD && N == 64) return launch_bwd64(qkv, out, dout, lse, nullptr, dqkv, B, H, scale, stream);
  if (small_ok(N, hd)) return dispatch_small(true, qkv, dout, dqkv, lse, nullptr, B, N, H, hd, scale, stream);
  int wpc, pw;
  TAE_CHECK_SHAPE(simt_config(N, hd, true, &wpc, &pw) == 0, "tae_attention_bwd: N=%d hd=%d does not fit shared memory", N, hd);
  const int smem = wpc * pw * 4;
  int rc = set_smem(attn_bwd_simt, smem);
  if (rc) return rc;
  const int BH = B * H;
  attn_bwd_simt<<<(BH + wpc - 1) / wpc, wpc * 32, smem, stream>>>(qkv, dout, lse, dqkv, BH, N, H, hd, scale, pw);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

extern "C" int tae_attention_bwd_delta(const tae_bf16* qkv_, const tae_bf16* dout_, const float* lse, const float* delta,
                                       tae_bf16* dqkv_, int32_t B, int32_t N, int32_t H, int32_t hd, void* stream_) {
  using namespace tae;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(qkv_ && dout_ && lse && delta && dqkv_, "tae_attention_bwd_delta: NULL argument");
  TAE_CHECK_SHAPE(B > 0 && H > 0 && (N == 256 || N == 64) && hd == 64,
                  "tae_attention_bwd_delta: only the N=256 / N=64, hd=64 kernels take a precomputed delta (got N=%d hd=%d)", N, hd);
  if (N == 64)
    return attn::launch_bwd64(reinterpret_cast<const bf16*>(qkv_), nullptr, reinterpret_cast<const bf16*>(dout_), lse, delta,
                              reinterpret_cast<bf16*>(dqkv_), B, H, 1.0f / sqrtf((float)hd), reinterpret_cast<cudaStream_t>(stream_));
  return attention_bwd_tcgen05(reinterpret_cast<const bf16*>(qkv_), nullptr, reinterpret_cast<const bf16*>(dout_), lse, delta,
                               reinterpret_cast<bf16*>(dqkv_), B, H, stream);
}
