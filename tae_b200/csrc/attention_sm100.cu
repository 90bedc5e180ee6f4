// attention_sm100.cu — whole-sequence attention on the tcgen05 tensor cores for the 256-token grid (patch16:
// N = 256, head_dim = 64), forward and backward.  Same contract as attention.cu (tae.py:74-80): packed qkv
// [B*N, 3*H*64] in, merged [B*N, H*64] out, log-sum-exp saved for backward.
//
// Forward (attn_fwd_tc_ring): TMA operands, S = Q K^T in 128x128 half tiles through a ring of TMEM buffers, two softmax
//   groups (exp2 domain, lazy reference across the key halves), P written back to TMEM as the A operand of O = P V
//   (V consumed MN-major exactly as it sits in the qkv buffer), O / l -> bf16 -> swizzled smem -> TMA store;
//   lse = m*ln2 + ln(l).
//
// Backward (attn_bwd_tc_persist): everything is computed transposed (keys on TMEM lanes) so that every product is a plain
// UMMA with operands that already exist in shared memory:
//   S^T = K Q^T, dP^T = V dO^T                               (K-major x K-major)
//   P^T = exp2(S^T*c - lse2[q]),  dS^T = P^T (dP^T - delta[q]) * scale      (thread-per-key-row, written as bf16 A tiles)
//   dV += P^T dO,  dK += dS^T Q                              (A K-major, B = dO / Q MN-major in place)
//   dQ += dS K                                               (A = the SAME dS^T tile read MN-major, B = K MN-major)
// in [128 keys x 64 queries] blocks with double-buffered S^T / dP^T accumulators.  No atomics, deterministic.
#include <stdlib.h>

#include "sm100.cuh"

namespace tae {
extern long long* g_attn_trace;
namespace attn_tc {

using namespace tae::sm100;

constexpr int N = 256;
constexpr int HD = 64;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// byte offset of 16-byte chunk `chunk` of row `row` inside a [rows x 128 B] 128B-swizzled tile
__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// =============================================================================================
// Forward: persistent, one CTA per SM walks (image, head) items, two items' operands resident (2 x 96 KB, TMA straight out
// of the packed qkv buffer).
//   warp 0   MMA issue          warp 1   TMA producer (K, Q, V of item i+1 as soon as its buffer is free)
//   warps 4-11 / 12-19   softmax + epilogue of query tile 0 / 1: two threads per row
// The scores are produced in HALF tiles (128 queries x 128 keys = 128 TMEM columns) that rotate through a ring of three
// buffers, and the two O accumulators have their own columns, so the tensor core always works one half tile ahead of
// each softmax group (a kernel whose S tile shared its columns with P and O ran the chain S -> max -> exp -> PV -> drain
// serially per item: 0.172 ms against 0.144 ms for this one at B=256, H=16):
//   TMEM  [0,384): ring of 3 score buffers (S half tile, then P over it: thread half h keeps its 64 keys' P in columns
//                  +64h .. +64h+31)          [384,512): O_0, O_1 (64 columns each)
//   jobs of item i, in order: (tile 0, keys 0-127), (tile 1, keys 0-127), (tile 0, keys 128-255), (tile 1, keys 128-255);
//   job g uses ring buffer g % 3.  MMA thread, iteration g: S(g+2) as soon as PV(g-1) has released its buffer, then
//   PV(g) as soon as the group has written P(g) (bf16, read by the MMA as a TMEM A operand).
//   Softmax across the two key halves without rescaling O: the second half keeps the FIRST half's row maximum as its
//   reference (softmax is shift-invariant; bf16 P and fp32 l, O only need the exponent range), unless the second half's
//   maximum exceeds it by more than 2^32 — only then O and l are rescaled (warp-uniform slow path).
//   A group's order of work per item: second half of item i, FIRST half of item i+1, then the epilogue of item i — the
//   PV of item i's second half finishes behind the next softmax instead of in front of an idle group.
// The kernel is bound by the exponentials (65536 MUFU.EX2 per item = 4.1k cycles per SM) and their issue slots; moving a
// quarter or half of them to a degree-3 polynomial on the FMA pipes was measured and is not faster (0.150 / 0.152 ms).
// =============================================================================================
constexpr int G_BUF = 98304;                 // one item's operands: Q 32 KB | K 32 KB | V 32 KB
constexpr int R_OFF_RED = 2 * G_BUF;         // [2 tiles][ max lo | max hi | sum ][2 halves][128] fp32 = 6 KB (8 KB reserved)
constexpr int R_OFF_BAR = R_OFF_RED + 8192;
constexpr int R_SMEM = R_OFF_BAR + 256 + 1024;
constexpr int R_THREADS = 640;
constexpr float R_TAU = 32.0f;               // log2 of the largest P the lazy reference may produce

__global__ void __launch_bounds__(R_THREADS, 1)
attn_fwd_tc_ring(const __grid_constant__ CUtensorMap tm_qkv,  // box 256 rows x 64 cols over qkv [B*N, 3D]
                 const __grid_constant__ CUtensorMap tm_o,    // box 128 rows x 64 cols over out [B*N, D]
                 float* __restrict__ lse, int H, int num_items, float scale, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R_OFF_BAR);
  uint64_t* bar_kq = bars;           // [2 buffers] K and Q landed                      (TMA, every other item)
  uint64_t* bar_v = bars + 2;        // [2 buffers] V landed
  uint64_t* bar_buffree = bars + 4;  // [2 buffers] both tiles' output stores have read the buffer (2 arrivals)
  uint64_t* bar_s = bars + 6;        // [3 ring]    S half tile in TMEM                 (commit, every 3rd job)
  uint64_t* bar_p = bars + 9;        // [3 ring]    P written over it                   (256 arrivals)
  uint64_t* bar_pv = bars + 12;      // [3 ring]    PV of that job done: buffer free, O updated (commit)
  uint64_t* bar_ofree = bars + 15;   // [2 tiles]   O_t read out of TMEM                (256 arrivals, every item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * HD;
  const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_kq[i], 1);
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_buffree[i], 2);
      mbar_init(&bar_ofree[i], 256);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 256);
      mbar_init(&bar_pv[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 1) {
    if (elect_one()) {
      // ---------------- TMA producer: K, Q, V of item i into operand buffer i & 1 ----------------
#pragma unroll 1
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int b = item / H, h = item - b * H;
        const int s = it & 1;
        if (it >= 2) mbar_wait(&bar_buffree[s], (uint32_t)(((it >> 1) - 1) & 1));
        uint8_t* buf = smem + s * G_BUF;
        mbar_expect_tx(&bar_kq[s], 65536);
        tma_load_2d(buf + 32768, &tm_qkv, &bar_kq[s], D + h * HD, b * N);
        tma_load_2d(buf, &tm_qkv, &bar_kq[s], h * HD, b * N);
        mbar_expect_tx(&bar_v[s], 32768);
        tma_load_2d(buf + 65536, &tm_qkv, &bar_v[s], 2 * D + h * HD, b * N);
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const int G = 4 * my_items;
      auto issue_s = [&](int g) {
        const int it = g >> 2, j = g & 3, t = j & 1, hi = j >> 1, s = it & 1, rb = g % 3;
        const uint32_t base = smem_u32(smem + s * G_BUF);
        mbar_wait(&bar_kq[s], (uint32_t)((it >> 1) & 1));
        if (g >= 3) mbar_wait(&bar_pv[rb], (uint32_t)(((g - 3) / 3) & 1));  // PV(g-3) has consumed the P in this buffer
        tcgen05_fence_after();
        const uint32_t q_lo = desc_lo(base + t * 16384), k_lo = desc_lo(base + 32768 + hi * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_lo(tmem + rb * 128, q_lo + k * 2, k_lo + k * 2, idesc_s, k > 0);
        umma_commit(&bar_s[rb]);
      };
      auto issue_pv = [&](int g) {
        const int it = g >> 2, j = g & 3, t = j & 1, hi = j >> 1, s = it & 1, rb = g % 3;
        const uint32_t sV = smem_u32(smem + s * G_BUF + 65536);
        mbar_wait(&bar_v[s], (uint32_t)((it >> 1) & 1));
        if (!hi && it >= 1) mbar_wait(&bar_ofree[t], (uint32_t)((it - 1) & 1));  // the previous item's O_t has been read
        mbar_wait(&bar_p[rb], (uint32_t)((g / 3) & 1));
        tcgen05_fence_after();
        const uint32_t o_t = tmem + 384 + t * 64, p_t = tmem + rb * 128;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)  // 16 keys = 8 TMEM columns of P per instruction
          umma_f16_ts(o_t, p_t + (jj >> 2) * 64 + (jj & 3) * 8, make_smem_desc(sV + (hi * 8 + jj) * 2048, 8192, 1024), idesc_pv,
                      (hi || jj > 0) ? 1u : 0u);
        umma_commit(&bar_pv[rb]);
      };
      if (G > 0) issue_s(0);
      if (G > 1) issue_s(1);
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        if (g + 2 < G) issue_s(g + 2);
        issue_pv(g);
      }
    }
  } else if (warp >= 4) {
    // ---------------- softmax + epilogue groups ----------------
    const int t = (warp - 4) >> 3;               // query tile of this group
    const int wg = (warp - 4) & 7;               // warp inside the group
    const int q = warp & 3;                      // TMEM lane quarter (== warp % 4)
    const int half = wg >> 2;                    // which 64 keys of a 128-key half tile
    const int r = q * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    float* red = reinterpret_cast<float*>(smem + R_OFF_RED) + t * 768;  // [max lo | max hi | sum][2 halves][128]
    const bool storer = (wg == 0);
    // m2: the row's softmax reference in the exp2 domain, an INTEGER (rint of max * scale * log2 e) — any reference
    // gives the same softmax, an integer one keeps MAGIC - m2 exact for the polynomial path and makes rescales powers of 2
    float m2 = 0.f, l = 0.f;

    // one half-tile job: scores of ring buffer g % 3 -> P (bf16, over the scores); updates m2 / l
    auto softmax_job = [&](int g, int hi) {
      const int rb = g % 3;
      const uint32_t tS = tlane + (uint32_t)(rb * 128 + half * 64);
      mbar_wait(&bar_s[rb], (uint32_t)((g / 3) & 1));
      tcgen05_fence_after();
      uint32_t buf[2][32];
      tmem_ld_32x32b_x32(tS, buf[0]);
      tmem_ld_32x32b_x32(tS + 32, buf[1]);
      tmem_ld_wait_regs(buf[0]);
      tmem_ld_wait_regs(buf[1]);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(buf[0][i]));
        mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(buf[1][i]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      float* s_max = red + hi * 256;
      s_max[half * 128 + r] = mx;
      tmem_ld_32x32b_x32(tS, buf[0]);  // first chunk of the exp pass in flight across the barrier
      named_bar_sync(1 + t, 256);
      mx = fmaxf(mx, s_max[(half ^ 1) * 128 + r]);  // row maximum over the 128 keys of this half tile
      if (!hi) {
        m2 = rintf(mx * sl2);
        l = 0.f;
      } else {
        // lazy reference: keep the first half's maximum unless this half's exceeds it by more than 2^R_TAU
        const bool need = fmaf(mx, sl2, -m2) > R_TAU;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? rintf(mx * sl2) : m2;
          const float f = ex2_approx(m2 - m_new);
          l *= f;
          m2 = m_new;
          // O_t holds the first half's product: job g - 2 of this tile
          mbar_wait(&bar_pv[(g - 2) % 3], (uint32_t)(((g - 2) / 3) & 1));
          tcgen05_fence_after();
          tmem_ld_wait_regs(buf[0]);  // (the chunk already in flight)
          uint32_t o[32];
          const uint32_t tO = tlane + (uint32_t)(384 + t * 64 + half * 32);
          tmem_ld_32x32b_x32(tO, o);
          tmem_ld_wait_regs(o);
          uint32_t o_lo[16], o_hi[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            o_lo[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            o_hi[i] = __float_as_uint(__uint_as_float(o[16 + i]) * f);
          }
          tmem_st_32x32b_x16(tO, o_lo);
          tmem_st_32x32b_x16(tO + 16, o_hi);
          tmem_st_wait();
        }
      }
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_wait_regs(buf[c]);
        if (c == 0) tmem_ld_32x32b_x32(tS + 32, buf[1]);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(buf[c][2 * i]), sl2, -m2));
          const float p1 = ex2_approx(fmaf(__uint_as_float(buf[c][2 * i + 1]), sl2, -m2));
          l0 += p0;
          l1 += p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x16(tS + c * 16, pk);  // P chunk c over score columns this thread has already consumed
      }
      l += l0 + l1;
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(&bar_p[rb]);
    };

    if (my_items > 0) softmax_job(t, 0);
#pragma unroll 1
    for (int it = 0; it < my_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = item / H, h = item - b * H;
      const int s = it & 1;
      const int g_hi = 4 * it + 2 + t;
      softmax_job(g_hi, 1);
      const float m2_fin = m2, l_part = l;
      if (it + 1 < my_items) softmax_job(4 * (it + 1) + t, 0);  // overwrites m2 / l with the next item's
      // ---- epilogue of item `it`: each thread normalises 32 of the row's 64 output columns ----
      float* s_sum = red + 512;
      s_sum[half * 128 + r] = l_part;
      named_bar_sync(1 + t, 256);
      const float l_tot = l_part + s_sum[(half ^ 1) * 128 + r];
      mbar_wait(&bar_pv[g_hi % 3], (uint32_t)((g_hi / 3) & 1));  // O_t complete
      tcgen05_fence_after();
      if (half == 0) lse[((size_t)b * H + h) * N + t * 128 + r] = fmaf(m2_fin, 0.69314718055994530942f, __logf(l_tot));
      const float inv = 1.0f / l_tot;
      const uint32_t stage = smem_u32(smem + s * G_BUF + t * 16384);  // the item's dead Q tile
      {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(tlane + (uint32_t)(384 + t * 64 + half * 32), raw);
        tmem_ld_wait_regs(raw);
        tcgen05_fence_before();
        mbar_arrive(&bar_ofree[t]);  // O_t may take the next item's first PV
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = pack_bf16x2(__uint_as_float(raw[i * 8 + 2 * j]) * inv, __uint_as_float(raw[i * 8 + 2 * j + 1]) * inv);
          st_shared_v4(stage + sw128(r, half * 4 + i), o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 256);  // staging complete; also: every thread has read s_sum before the next item rewrites it
      if (storer && elect_one()) {
        tma_store_2d(&tm_o, smem + s * G_BUF + t * 16384, h * HD, b * N + t * 128);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&bar_buffree[s]);  // (with the other tile's arrival) the producer may refill this buffer
      }
    }
    if (storer && elect_one()) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// Backward
// =============================================================================================
constexpr int P_OFF_Q = 0, P_OFF_K = 32768, P_OFF_V = 65536, P_OFF_DO = 98304;
constexpr int P_OFF_PT = 131072;     // 2 x [128 x 64] bf16 (16 KB each)
constexpr int P_OFF_DST = 163840;    // 2 x [128 x 64] bf16
constexpr int PC_S = 0, PC_DP = 128, PC_DV = 256, PC_DK = 320, PC_DQ = 384;  // TMEM column map

__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

// =============================================================================================
// Backward, PERSISTENT variant (default).  One CTA per SM walks a static list of (image, head) items; the block
// pipeline of attn_bwd_tc_pipe runs unchanged INSIDE an item, and three things overlap ACROSS items:
//   * operand prefetch: Q/dO query blocks are re-loaded for the next item as soon as the last MMA that reads them has
//     retired (during the second key-tile pass), K0/V0 once their dV0/dK0 staging stores have left shared memory, the
//     rest right after the final MMA — so the next item's first scores are in TMEM before the element-wise warps get
//     there (the 128 KB operand load, ~8k cycles exposed per item in the one-shot kernel, disappears);
//   * delta = rowsum(dO*O) and lse of the NEXT item are prepared by three otherwise idle warps into a second buffer;
//   * the final accumulator drain goes through the (dead) P^T/dS^T tiles, so the operand tiles are free for the
//     prefetch while the results are still on their way out through TMA stores.
// Shared memory map = attn_bwd_tc_pipe's, with lse/delta double-buffered.
// =============================================================================================
constexpr int S_OFF_Q0ALT = 196608;   // second home of Q block 0 (odd items), 8 KB
constexpr int S_OFF_DO0ALT = 204800;  // second home of dO block 0 (odd items), 8 KB
constexpr int S_OFF_LSE = 212992;     // [2][256] fp32
constexpr int S_OFF_DELTA = 215040;   // [2][256] fp32
constexpr int S_OFF_BAR = 217088;
constexpr int S_SMEM = S_OFF_BAR + 256 + 1024;
constexpr int S_THREADS = 512;  // warp 0: MMA issue; 1: TMA producer; 2-3: delta/lse of the next item; 4-11: element-wise; 12-15: drains

__global__ void __launch_bounds__(S_THREADS, 1)
attn_bwd_tc_persist(const __grid_constant__ CUtensorMap tm_qkv,   // box 64 rows x 64 cols over qkv  [B*N, 3D]
                    const __grid_constant__ CUtensorMap tm_do,    // box 64 rows x 64 cols over dout [B*N, D]
                    const __grid_constant__ CUtensorMap tm_dqkv,  // box 128 rows x 64 cols over dqkv (stores)
                    const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta_in,  // optional [B, H, N]: rowsum(dO * O) from the GEMM that made dO
                    int H, int num_items, float scale, float sl2, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_OFF_BAR);
  // debug timeline (tools/attn_trace.py): CTA gridDim.x/2 stamps clock64() at the milestones of its 6th item
  const bool trc = trace != nullptr && blockIdx.x == gridDim.x / 2;
#define TRACE_C(i) do { if (trc && it == 5) trace[(i)] = clock64(); } while (0)
#define TRACE_E(i) do { if (trc && it == 5 && threadIdx.x == 128) trace[32 + (i)] = clock64(); } while (0)
  uint64_t* bK = bars;               // [2] K tile kt landed                        (TMA, once per item)
  uint64_t* bV = bars + 2;           // [2] V tile kt landed
  uint64_t* bQ = bars + 4;           // [4] Q block j landed
  uint64_t* bdO = bars + 8;          // [4] dO block j landed
  uint64_t* bar_s = bars + 12;       // [2] S^T/dP^T buffer filled                  (4 per item each)
  uint64_t* bar_p = bars + 14;       // [2] P^T/dS^T tile written, S/dP drained     (256 arrivals)
  uint64_t* bar_pfree = bars + 16;   // [2] gradient MMAs reading the tile retired
  uint64_t* bar_g = bars + 18;       // dV/dK (at the end dQ) complete              (2 per item)
  uint64_t* bar_dfree = bars + 19;   // dV0/dK0 left TMEM                           (128 drain threads, once per item)
  uint64_t* bar_accfree = bars + 20; // dV1/dK1/dQ left TMEM                        (128, once per item)
  uint64_t* bar_kv0free = bars + 21; // dV0/dK0 staging stores have read K0/V0      (1, once per item)
  uint64_t* bar_tailfree = bars + 27;// final staging stores have read K1/V1/Q2-3/dO2-3 (1, once per item)
  uint64_t* bQalt = bars + 28;       // Q block 0 landed in its second home            (TMA, every other item)
  uint64_t* bdOalt = bars + 29;      // dO block 0 landed in its second home
  uint64_t* bar_qfree = bars + 30;   // [1] is used: the last MMA reading Q/dO block 1 of the item retired (commit, once per item)
  uint64_t* bar_delta = bars + 22;   // [2] lse/delta buffer filled                 (64, every other item each)
  uint64_t* bar_dbuffree = bars + 24;// [2] lse/delta buffer consumed               (256, every other item each)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);  // (bars + 27 is bar_tailfree)
  float* sLse = reinterpret_cast<float*>(smem + S_OFF_LSE);
  float* sDelta = reinterpret_cast<float*>(smem + S_OFF_DELTA);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = H * HD;
  // this CTA's items: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqkv);
    for (int i = 0; i < 12; ++i) mbar_init(&bars[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], 256);
      mbar_init(&bar_pfree[i], 1);
      mbar_init(&bar_delta[i], 64);
      mbar_init(&bar_dbuffree[i], 256);
    }
    mbar_init(bar_g, 1);
    mbar_init(bar_dfree, 128);
    mbar_init(bar_accfree, 128);
    mbar_init(bar_kv0free, 1);
    mbar_init(bar_tailfree, 1);
    mbar_init(bQalt, 1);
    mbar_init(bdOalt, 1);
    mbar_init(&bar_qfree[0], 1);
    mbar_init(&bar_qfree[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + P_OFF_Q), sK = smem_u32(smem + P_OFF_K), sV = smem_u32(smem + P_OFF_V);
  const uint32_t sdO = smem_u32(smem + P_OFF_DO), sPT = smem_u32(smem + P_OFF_PT), sdST = smem_u32(smem + P_OFF_DST);

  if (warp == 1) {
    if (elect_one()) {
      // ---------------- TMA producer: every operand piece of item it+1 as soon as its home is dead ----------------
      auto item_bh = [&](int it, int& b, int& h) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        b = item / H;
        h = item - b * H;
      };
      auto load_q = [&](int it, int j) {  // Q block j and dO block j of item `it` (block 0 alternates between two homes)
        int b, h;
        item_bh(it, b, h);
        // each home of block 0 has its own barrier (it completes every other item); blocks 1-3 complete every item
        const bool alt = (j == 0) && (it & 1);
        uint64_t* bq = alt ? bQalt : &bQ[j];
        uint64_t* bo = alt ? bdOalt : &bdO[j];
        mbar_expect_tx(bq, 8192);
        tma_load_2d(smem + (alt ? S_OFF_Q0ALT : P_OFF_Q + j * 8192), &tm_qkv, bq, h * HD, b * N + j * 64);
        mbar_expect_tx(bo, 8192);
        tma_load_2d(smem + (alt ? S_OFF_DO0ALT : P_OFF_DO + j * 8192), &tm_do, bo, h * HD, b * N + j * 64);
      };
      auto load_kv = [&](int it, int kt) {
        int b, h;
        item_bh(it, b, h);
        mbar_expect_tx(&bK[kt], 16384);
        tma_load_2d(smem + P_OFF_K + kt * 16384, &tm_qkv, &bK[kt], D + h * HD, b * N + kt * 128);
        tma_load_2d(smem + P_OFF_K + kt * 16384 + 8192, &tm_qkv, &bK[kt], D + h * HD, b * N + kt * 128 + 64);
        mbar_expect_tx(&bV[kt], 16384);
        tma_load_2d(smem + P_OFF_V + kt * 16384, &tm_qkv, &bV[kt], 2 * D + h * HD, b * N + kt * 128);
        tma_load_2d(smem + P_OFF_V + kt * 16384 + 8192, &tm_qkv, &bV[kt], 2 * D + h * HD, b * N + kt * 128 + 64);
      };
      if (my_items > 0) {
        load_kv(0, 0);
        load_q(0, 0);
        load_q(0, 1);
        load_q(0, 2);
        load_q(0, 3);
        load_kv(0, 1);
      }
#pragma unroll 1
      for (int it = 0; it + 1 < my_items; ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        // Q/dO block 0 goes to its OTHER home, last read by item it-1: all of that item had retired when its
        // bar_tailfree (waited at the end of the previous iteration) completed
        load_q(it + 1, 0);
        // K0/V0: the drain warps' dV0/dK0 staging stores of this item have read them
        mbar_wait(bar_kv0free, par);
        load_kv(it + 1, 0);
        // Q/dO block 1: its last reader of this item (block 5) has retired
        mbar_wait(&bar_qfree[1], par);
        load_q(it + 1, 1);
        // K1/V1 and Q/dO blocks 2-3 are the drain warps' final staging tiles until their stores have read them
        mbar_wait(bar_tailfree, par);
        load_q(it + 1, 2);
        load_q(it + 1, 3);
        load_kv(it + 1, 1);
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t id_s = make_idesc_bf16(128, 64, 0, 0);    // S^T, dP^T: [128 keys x 64 q]
      const uint32_t id_kn = make_idesc_bf16(128, 64, 0, 1);   // dV, dK
      const uint32_t id_q = make_idesc_bf16(64, 64, 1, 1);     // dQ: M = 64 queries
      const uint32_t q_lo = desc_lo(sQ), k_lo = desc_lo(sK), v_lo = desc_lo(sV), do_lo = desc_lo(sdO);
      const uint32_t pt_lo = desc_lo(sPT), dst_lo = desc_lo(sdST);
      const uint32_t qmn_lo = desc_lo(sQ, 8192), kmn_lo = desc_lo(sK, 8192), domn_lo = desc_lo(sdO, 8192);
      const uint32_t dstmn_lo = desc_lo(sdST, 8192);
      const uint32_t sQalt = smem_u32(smem + S_OFF_Q0ALT), sdOalt = smem_u32(smem + S_OFF_DO0ALT);
      const uint32_t qalt_lo = desc_lo(sQalt), doalt_lo = desc_lo(sdOalt);
      const uint32_t qaltmn_lo = desc_lo(sQalt, 8192), doaltmn_lo = desc_lo(sdOalt, 8192);
      // scores of block `blk` of item `it`: S^T = K_kt Q_j^T, dP^T = V_kt dO_j^T into S/dP buffer (blk & 1)
      auto issue_scores = [&](int it, int blk) {
        const int kt = blk >> 2, j = blk & 3, buf = blk & 1;
        const uint32_t par = (uint32_t)(it & 1);
        const uint32_t ka = k_lo + kt * (16384 >> 4), va = v_lo + kt * (16384 >> 4);
        const bool alt = (j == 0) && (it & 1);
        const uint32_t qb = alt ? qalt_lo : q_lo + j * (8192 >> 4), ob = alt ? doalt_lo : do_lo + j * (8192 >> 4);
        const uint32_t ds = tmem + PC_S + buf * 64, dp = tmem + PC_DP + buf * 64;
        const uint32_t parq = j == 0 ? (uint32_t)((it >> 1) & 1) : par;
        mbar_wait(&bK[kt], par);
        mbar_wait(alt ? bQalt : &bQ[j], parq);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_lo(ds, ka + k * 2, qb + k * 2, id_s, k > 0);
        mbar_wait(&bV[kt], par);
        mbar_wait(alt ? bdOalt : &bdO[j], parq);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_lo(dp, va + k * 2, ob + k * 2, id_s, k > 0);
        umma_commit(&bar_s[buf]);
      };
      if (my_items > 0) {
        issue_scores(0, 0);
        issue_scores(0, 1);
      }
#pragma unroll 1
      for (int it = 0; it < my_items; ++it) {
        const bool has_next = it + 1 < my_items;
#pragma unroll 1
        for (int blk = 0; blk < 8; ++blk) {
          const int kt = blk >> 2, j = blk & 3, buf = blk & 1;
          const int g = it * 8 + blk;  // running block index: barrier phases are counted over the whole CTA lifetime
          mbar_wait(&bar_p[buf], (uint32_t)((g >> 1) & 1));
          tcgen05_fence_after();
          TRACE_C(1 + 2 * blk);
          if (blk == 0 && it > 0) {  // dV1/dK1/dQ of the previous item must have left TMEM before they are overwritten
            mbar_wait(bar_accfree, (uint32_t)((it - 1) & 1));
            tcgen05_fence_after();
          }
          if (blk == 4) {  // dV0/dK0 must have left TMEM
            mbar_wait(bar_dfree, (uint32_t)(it & 1));
            tcgen05_fence_after();
          }
          const uint32_t a_pt = pt_lo + buf * (16384 >> 4), a_dst = dst_lo + buf * (16384 >> 4);
          const bool alt = (j == 0) && (it & 1);
          const uint32_t b_do = alt ? doaltmn_lo : domn_lo + j * (8192 >> 4);
          const uint32_t b_q = alt ? qaltmn_lo : qmn_lo + j * (8192 >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 queries per instruction
            umma_f16_lo(tmem + PC_DV, a_pt + k * 2, b_do + k * (2048 >> 4), id_kn, (j > 0 || k > 0));
            umma_f16_lo(tmem + PC_DK, a_dst + k * 2, b_q + k * (2048 >> 4), id_kn, (j > 0 || k > 0));
          }
          const uint32_t dq_addr = tmem + PC_DQ + (j >> 1) * 64 + ((uint32_t)((j & 1) * 16) << 16);
          const uint32_t a_ds = dstmn_lo + buf * (16384 >> 4), b_k = kmn_lo + kt * (16384 >> 4);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 keys per instruction; A = dS^T tile read MN-major (64 queries contiguous)
            umma_f16_lo(dq_addr, a_ds + k * (2048 >> 4), b_k + k * (2048 >> 4), id_q, (kt > 0 || k > 0));
          umma_commit(&bar_pfree[buf]);
          if (j == 3) umma_commit(bar_g);
          if (blk == 5) umma_commit(&bar_qfree[1]);  // block 5 = (kt 1, j 1) is the last reader of Q/dO block 1
          // scores run two blocks ahead, across the item boundary (S/dP buffer blk&1 was drained by block blk)
          if (blk < 6) {
            issue_scores(it, blk + 2);
          } else if (has_next) {
            issue_scores(it + 1, blk - 6);
          }
          TRACE_C(2 + 2 * blk);
        }
      }
    }
  } else if (warp < 4) {
    // ---------------- helper warps 2-3: delta = rowsum(dO * O) and lse (exp2 domain) of item `it` ----------------
    const int te = threadIdx.x - 64;  // 0..63
#pragma unroll 1
    for (int it = 0; it < my_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = item / H, h = item - b * H;
      const int dbuf = it & 1;
      if (it >= 2) mbar_wait(&bar_dbuffree[dbuf], (uint32_t)(((it >> 1) - 1) & 1));
      const bf16* go = out + (size_t)b * N * D + (size_t)h * HD;
      const bf16* gdo = dout + (size_t)b * N * D + (size_t)h * HD;
      float* dl = sDelta + dbuf * 256;
      float* ls = sLse + dbuf * 256;
      if (delta_in != nullptr) {
        // delta came with dO: two 1 KB rows to copy per item
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(delta_in + ((size_t)b * H + h) * N) + te);
        const float4 l4 = __ldg(reinterpret_cast<const float4*>(lse + ((size_t)b * H + h) * N) + te);
        constexpr float kL2e = 1.44269504088896340736f;
        reinterpret_cast<float4*>(dl)[te] = d4;
        reinterpret_cast<float4*>(ls)[te] = make_float4(l4.x * kL2e, l4.y * kL2e, l4.z * kL2e, l4.w * kL2e);
        mbar_arrive(&bar_delta[dbuf]);
        continue;
      }
      // 4 chunks of 64 rows; all 16 loads of a chunk are in flight before the first use (4 round trips per item)
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 64) {
        uint4 dv[8], ov[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int row = c0 + u * 8 + (te >> 3), ch = te & 7;
          dv[u] = make_uint4(0u, 0u, 0u, 0u);
          ov[u] = make_uint4(0u, 0u, 0u, 0u);
          if (row < 256) {
            dv[u] = ld_nc_v4(gdo + (size_t)row * D + ch * 8);
            ov[u] = ld_nc_v4(go + (size_t)row * D + ch * 8);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int row = c0 + u * 8 + (te >> 3), ch = te & 7;
          const uint32_t dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w}, ow[4] = {ov[u].x, ov[u].y, ov[u].z, ov[u].w};
          float acc = 0.f;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float2 d2 = unpack_bf16x2(dw[jj]), o2 = unpack_bf16x2(ow[jj]);
            acc += d2.x * o2.x + d2.y * o2.y;
          }
          acc += __shfl_xor_sync(0xffffffffu, acc, 1);
          acc += __shfl_xor_sync(0xffffffffu, acc, 2);
          acc += __shfl_xor_sync(0xffffffffu, acc, 4);
          if (ch == 0 && row < 256) dl[row] = acc;
        }
      }
      for (int i = te; i < 256; i += 64) ls[i] = lse[((size_t)b * H + h) * N + i] * 1.44269504088896340736f;
      mbar_arrive(&bar_delta[dbuf]);  // release: the smem writes above are visible to whoever acquires the phase
    }
  } else if (warp < 12) {
    // ---------------- element-wise warps 4-11 ----------------
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;   // which 32 query columns of the 64-wide block
    const int r = q4 * 32 + lane;       // key row inside the key tile == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
    for (int it = 0; it < my_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = item / H, h = item - b * H;
      const int dbuf = it & 1;
      TRACE_E(29);
      mbar_wait(&bar_delta[dbuf], (uint32_t)((it >> 1) & 1));
      TRACE_E(0);
      const uint32_t sLseA = smem_u32(sLse + dbuf * 256), sDeltaA = smem_u32(sDelta + dbuf * 256);
#pragma unroll 1
      for (int blk = 0; blk < 8; ++blk) {
        const int kt = blk >> 2, j = blk & 3, buf = blk & 1;
        const int g = it * 8 + blk;
        mbar_wait(&bar_s[buf], (uint32_t)((g >> 1) & 1));
        tcgen05_fence_after();
        TRACE_E(1 + 3 * blk);
        uint32_t sraw[32], draw[32];
        tmem_ld_32x32b_x32(tlane + PC_S + buf * 64 + half * 32, sraw);
        tmem_ld_32x32b_x32(tlane + PC_DP + buf * 64 + half * 32, draw);
        tmem_ld_wait();
        const int qcol = j * 64 + half * 32;
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 lv = ld_shared_f4(sLseA + (qcol + 4 * i) * 4), dl = ld_shared_f4(sDeltaA + (qcol + 4 * i) * 4);
          const float p0 = ex2_approx(fmaf(__uint_as_float(sraw[4 * i + 0]), sl2, -lv.x));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sraw[4 * i + 1]), sl2, -lv.y));
          const float p2 = ex2_approx(fmaf(__uint_as_float(sraw[4 * i + 2]), sl2, -lv.z));
          const float p3 = ex2_approx(fmaf(__uint_as_float(sraw[4 * i + 3]), sl2, -lv.w));
          pk[2 * i] = pack_bf16x2(p0, p1);
          pk[2 * i + 1] = pack_bf16x2(p2, p3);
          dk[2 * i] = pack_bf16x2(p0 * (__uint_as_float(draw[4 * i + 0]) - dl.x) * scale,
                                  p1 * (__uint_as_float(draw[4 * i + 1]) - dl.y) * scale);
          dk[2 * i + 1] = pack_bf16x2(p2 * (__uint_as_float(draw[4 * i + 2]) - dl.z) * scale,
                                      p3 * (__uint_as_float(draw[4 * i + 3]) - dl.w) * scale);
        }
        TRACE_E(2 + 3 * blk);
        if (g >= 2) mbar_wait(&bar_pfree[buf], (uint32_t)(((g >> 1) - 1) & 1));  // MMAs of block g-2 are done with this tile
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = (uint32_t)buf * 16384u + sw128(r, half * 4 + i);
          st_shared_v4(sPT + off, pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          st_shared_v4(sdST + off, dk[4 * i], dk[4 * i + 1], dk[4 * i + 2], dk[4 * i + 3]);
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        mbar_arrive(&bar_p[buf]);
        TRACE_E(3 + 3 * blk);
        if (blk == 7) mbar_arrive(&bar_dbuffree[dbuf]);  // last read of this item's lse/delta buffer
      }
    }
  } else {
    // ---------------- drain warps 12-15: accumulators -> bf16 -> swizzled staging tile -> TMA store ----------------
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;       // TMEM lane == key row inside the key tile
    const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
    // TMEM accumulator row (64 fp32 columns at `taddr`) -> bf16 -> row `row` of a 128B-swizzled [rows x 64] staging tile
    auto stage_row = [&](uint32_t taddr, uint32_t tile, int row) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(taddr + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_shared_v4(tile + sw128(row, c * 4 + i),
                       pack_bf16x2(__uint_as_float(raw[i * 8 + 0]), __uint_as_float(raw[i * 8 + 1])),
                       pack_bf16x2(__uint_as_float(raw[i * 8 + 2]), __uint_as_float(raw[i * 8 + 3])),
                       pack_bf16x2(__uint_as_float(raw[i * 8 + 4]), __uint_as_float(raw[i * 8 + 5])),
                       pack_bf16x2(__uint_as_float(raw[i * 8 + 6]), __uint_as_float(raw[i * 8 + 7])));
      }
    };
#pragma unroll 1
    for (int it = 0; it < my_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = item / H, h = item - b * H;
      // key tile 0: every MMA that reads K0 / V0 has retired (bar_g), so dK0 / dV0 are staged over those dead tiles
      mbar_wait(bar_g, (uint32_t)((it * 2) & 1));
      tcgen05_fence_after();
      stage_row(tlane + PC_DV, sV, r);
      stage_row(tlane + PC_DK, sK, r);
      tcgen05_fence_before();
      mbar_arrive(bar_dfree);
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (warp == 12 && elect_one()) {
        tma_store_2d(&tm_dqkv, smem + P_OFF_V, 2 * D + h * HD, b * N);
        tma_store_2d(&tm_dqkv, smem + P_OFF_K, D + h * HD, b * N);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(bar_kv0free);  // the producer may overwrite K0/V0 with the next item's tiles
      }
      // end of the item: dV1 -> V1, dK1 -> K1, dQ -> the Q / dO blocks 2-3 (two [128 x 64] tiles), all dead by now
      mbar_wait(bar_g, (uint32_t)((it * 2 + 1) & 1));
      tcgen05_fence_after();
      stage_row(tlane + PC_DV, sV + 16384, r);
      stage_row(tlane + PC_DK, sK + 16384, r);
      // dQ: M=64 accumulators; (quarter q4, lane l) of column group c2 owns query (2*c2 + (l>>4))*64 + 16*q4 + (l&15)
      const int qlo = (lane >> 4) * 64 + 16 * q4 + (lane & 15);  // row inside a 128-query tile
      stage_row(tlane + PC_DQ, sQ + 16384, qlo);
      stage_row(tlane + PC_DQ + 64, sdO + 16384, qlo);
      tcgen05_fence_before();
      mbar_arrive(bar_accfree);
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (warp == 12 && elect_one()) {
        tma_store_2d(&tm_dqkv, smem + P_OFF_V + 16384, 2 * D + h * HD, b * N + 128);
        tma_store_2d(&tm_dqkv, smem + P_OFF_K + 16384, D + h * HD, b * N + 128);
        tma_store_2d(&tm_dqkv, smem + P_OFF_Q + 16384, h * HD, b * N);
        tma_store_2d(&tm_dqkv, smem + P_OFF_DO + 16384, h * HD, b * N + 128);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(bar_tailfree);
      }
    }
    if (warp == 12 && elect_one()) tma_store_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 512);
#undef TRACE_C
#undef TRACE_E
}

}  // namespace attn_tc
long long* g_attn_trace = nullptr;  // debug: device buffer of 64 int64 set through tae_debug_set_attn_trace
namespace attn_tc {

template <typename K>
static int set_smem_once(K kernel, int bytes, cudaError_t* cached, std::once_flag* once) {
  std::call_once(*once, [&]() { *cached = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); });
  if (*cached != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%d) failed: %s", bytes, cudaGetErrorString(*cached));
    return TAE_ERR_CUDA;
  }
  return TAE_OK;
}

}  // namespace attn_tc

// entry points used by attention.cu's dispatcher (N == 256, hd == 64)
int attention_fwd_tcgen05(const bf16* qkv, bf16* out, float* lse, int B, int H, cudaStream_t stream) {
  using namespace attn_tc;
  const int D = H * HD;
  const float scale = 1.0f / sqrtf((float)HD);
  static cudaError_t err = cudaSuccess;
  static std::once_flag once;
  int rc = set_smem_once(attn_fwd_tc_ring, R_SMEM, &err, &once);
  if (rc) return rc;
  CUtensorMap tkv, to;
  rc = sm100::make_tmap(&tkv, qkv, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D, 256);
  if (rc) return rc;
  rc = sm100::make_tmap(&to, out, (uint64_t)D, (uint64_t)B * N, (uint64_t)D, 128);
  if (rc) return rc;
  const int sms = num_sms();
  if (sms <= 0) return TAE_ERR_CUDA;
  const int items = B * H;
  attn_fwd_tc_ring<<<items < sms ? items : sms, R_THREADS, R_SMEM, stream>>>(tkv, to, lse, H, items, scale,
                                                                             scale * 1.44269504088896340736f);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

int attention_bwd_tcgen05(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, const float* delta,
                          bf16* dqkv, int B, int H, cudaStream_t stream) {
  using namespace attn_tc;
  static cudaError_t err = cudaSuccess;
  static std::once_flag once;
  int rc = set_smem_once(attn_bwd_tc_persist, S_SMEM, &err, &once);
  if (rc) return rc;
  const int D = H * HD;
  const float scale = 1.0f / sqrtf((float)HD);
  const float sl2 = scale * 1.44269504088896340736f;
  CUtensorMap tq64, tdo64, tdq;
  rc = sm100::make_tmap(&tq64, qkv, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D, 64);
  if (rc) return rc;
  rc = sm100::make_tmap(&tdo64, dout, (uint64_t)D, (uint64_t)B * N, (uint64_t)D, 64);
  if (rc) return rc;
  rc = sm100::make_tmap(&tdq, dqkv, (uint64_t)3 * D, (uint64_t)B * N, (uint64_t)3 * D, 128);
  if (rc) return rc;
  const int sms = num_sms();
  if (sms <= 0) return TAE_ERR_CUDA;
  const int items = B * H;
  const int grid = items < sms ? items : sms;
  attn_bwd_tc_persist<<<grid, S_THREADS, S_SMEM, stream>>>(tq64, tdo64, tdq, out, dout, lse, delta, H, items, scale, sl2,
                                                            g_attn_trace);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

}  // namespace tae

// debug hook (not part of the public ABI): device buffer of 64 int64 receiving a clock64() timeline of one CTA
extern "C" void tae_debug_set_attn_trace(void* buf) { tae::g_attn_trace = reinterpret_cast<long long*>(buf); }
