// gemm_sm100.cu — persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (tcgen05) of sm_100a.
//
//   D[M,N] = sum_k A(m,k) * B(n,k)     fp32 accumulation in tensor memory (TMEM)
//
// Replaces every nn.Linear call site of the reference's hot path and its autograd dgrad/wgrad
// (tae.py:50,74,81,101,104,237,242,253) — see include/tae_b200.h for the operand conventions.
//
// Structure (one CTA per SM, 384 threads, static round-robin tile scheduler):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes (128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::f16 (128x256x16 per instruction)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered)
//   warps 4-11  epilogue: tcgen05.ld (one accumulator row per thread) -> swizzled smem transpose -> fused epilogue
//               with 128-byte coalesced global loads/stores
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and TMEM full/empty mbarriers (MMA <-> epilogue), so
// the epilogue of tile i overlaps the main loop of tile i+1.
//
// Operand majors: both operands may be K-major (k contiguous; forward) or MN-major (m/n contiguous; the
// transposed operands of dgrad/wgrad) — handled by the TMA box shape + UMMA smem-descriptor strides, so no
// transposed copy of any activation or weight is ever materialised.
#include <stdlib.h>

#include "sm100.cuh"

// Build-time switches of the CTA-pair kernel's epilogues (A/B variants: `python -m tae_b200.build --variant NAME -D...`).
// TAE_GELU_TMA_EPI (default 1): the GELU epilogue is evaluated in the accumulator's ROW layout, the layout tcgen05.ld
//   delivers, and leaves through TMA stores of swizzled bf16 staging boxes, instead of the fp32 transpose through shared
//   memory + coalesced st.global of the generic epilogue.  With it 8 epilogue warps carry the GELU work, which frees
//   the shared memory for a 6th pipeline stage.  Measured on B200 (fc1 of patch16, M=65536 N=4096 K=1024): generic
//   epilogue with 16 warps 0.488 ms, with 8 warps 0.531 ms; row-layout/TMA with 16 warps 0.494 ms, with 8 warps
//   0.465 ms; in the training step 998 -> 1092 TFLOP/s.  Dropping the second output (diagnostic) gives 0.43-0.44 ms:
//   the two output streams, not the GELU arithmetic, are what separates this GEMM from the plain one (0.39 ms).
// TAE_GELU_TMA_SWZ64: 64-byte swizzle of the staging boxes (bank-conflict-free st.shared); 0 = linear boxes.
// TAE_GELU_EW: epilogue warps for the GELU epilogue, 8 (6 pipeline stages) or 16 (5 stages).
// (The same path for the plain bf16(+bias) epilogue was parity-green but not faster — in-step 1311 vs 1319 TFLOP/s — and was removed.)
#ifndef TAE_GELU_TMA_EPI
#define TAE_GELU_TMA_EPI 1
#endif
#ifndef TAE_GELU_TMA_SWZ64
#define TAE_GELU_TMA_SWZ64 1
#endif
#ifndef TAE_GELU_EW
#define TAE_GELU_EW 8
#endif
// TAE_DGELU_TMA_EPI (default 1): row-layout GELU' epilogue — the gelu'(h) tile arrives by TMA load into the staging
//   box, the product overwrites it in place and leaves by TMA store, the bias column sums are read back from the staged
//   box; 8 epilogue warps, 6 stages.  fc2 dgrad of patch16 (M=65536 N=4096 K=1024): 0.478 -> 0.428 ms, in the training
//   step 1101 -> 1204 TFLOP/s.
#ifndef TAE_DGELU_TMA_EPI
#define TAE_DGELU_TMA_EPI 1
#endif
// The row-dot epilogue (proj dgrad emitting the attention backward's delta) runs on the same row-layout path: aux tile by TMA
//   load, bf16(acc + bias) written over it in place, TMA store; the per-head dot product is thread-local in the row
//   layout (no shuffles).  Measured in the patch16 step: 987 -> 1100-1120 TFLOP/s for that GEMM family
//   (profiles/r2_gemm_epilogue_ab.md).  The fp32 residual epilogue was tried on this path as well (16-column fp32 boxes by
//   TMA load / store): no gain (proj forward 0.150 ms either way — it sits on the HBM roof), so it keeps the generic path.

namespace tae {
namespace gemm {

using namespace tae::sm100;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B span
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int MN_BOX_BYTES = 64 * BLOCK_K * 2;        // one MN-major box: 64 (mn) x 64 (k) bf16 = 8 KB
constexpr int NUM_ACC = 2;                            // TMEM accumulator double buffer
constexpr int TMEM_COLS = NUM_ACC * BLOCK_N;          // 512
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;  // 384
constexpr int SMEM_BARRIER_BYTES = 512;
constexpr int SMEM_STAGING_BYTES = NUM_EPI_WARPS * 32 * 128;  // 4 KB transpose tile per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + SMEM_BARRIER_BYTES + SMEM_STAGING_BYTES + 1024;  // +1024: alignment slack

struct Params {
  int M, N, K;
  int a_mn, b_mn;
  int m_tiles, n_tiles, kb_total, kb_per_split, splits;
  int tile_n;          // CTA-pair kernel: output tile width (a multiple of 32, <= BLOCK_N; BLOCK_N unless chosen against wave quantisation)
  void* out;
  int ldo;
  void* out2;
  const float* bias;
  const float* resid;
  int ldr;
  int resid_rows;
  const bf16* aux;
  int ldaux;
  int beta;
  float* colsum_part;  // optional [ceil(M/32), N] fp32: per-32-row column sums of the bf16 output (bias gradients)
  float* rowdot;       // TAE_EPI_BF16_ROWDOT: fp32 [M / rd_tokens, N / 64, rd_tokens]
  int rd_tokens;
  int* sched;          // dynamic tile scheduler: {next work item, clusters done}; NULL = static round-robin
};

// K-major tile [rows x 64] (128 B per row, 8-row swizzle atoms of 1024 B): SBO = 1024 between 8-row groups;
// one UMMA_K=16 slice = 32 B inside the swizzle span.
__device__ __forceinline__ uint64_t desc_k_major(uint32_t tile_base, int k16) {
  return make_smem_desc(tile_base + (uint32_t)k16 * 32u, 0u, 1024u);
}
// MN-major tile stored as consecutive boxes of [64 k-rows x 64 mn] (128 B per k-row): atoms are 8 k-rows
// (SBO = 1024 between k-groups), LBO = 8192 between 64-wide mn chunks; one UMMA_K=16 slice = 2 k-groups = 2048 B.
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t tile_base, int k16) {
  return make_smem_desc(tile_base + (uint32_t)k16 * 2048u, (uint32_t)MN_BOX_BYTES, 1024u);
}
// Instruction descriptor for kind::f16: c=F32 (bit 4), a=b=BF16 (bits 7,10), majors (bits 15,16),
// N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Epilogue.  tcgen05.ld hands each thread one accumulator ROW (32 consecutive columns per load); writing global
// memory in that layout would touch 32 different 128-byte lines per instruction.  Each epilogue warp therefore
// transposes through a private 4 KB shared-memory tile (32 rows x 128 B, 16-byte chunks XOR-swizzled by row so both
// the row-wise writes and the line-wise reads are bank-conflict free) and then runs the fused epilogue in a
// COALESCED layout: 8 consecutive lanes own the 8 16-byte chunks of one 128-byte row segment, 4 rows per instruction.
// One step covers 32 accumulator columns; bias, bf16 rounding, GELU, residual all happen in the coalesced layout.
// ---------------------------------------------------------------------------------------------
constexpr int STG_BYTES_PER_WARP = 32 * 128;

__device__ __forceinline__ uint32_t stg_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}

// [32 rows x 32 bf16] TMA-store staging box (64 B per row); with CU_TENSOR_MAP_SWIZZLE_64B the 16-byte chunk index is
// XORed with address bits 7-8, i.e. (row >> 1) & 3 for 64-byte rows
__device__ __forceinline__ uint32_t stg64_off(int row, int chunk) {
#if TAE_GELU_TMA_SWZ64
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
#else
  return (uint32_t)(row * 64 + (chunk << 4));
#endif
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// TMA load into this CTA's shared memory (addresses given as shared-space integers), completing on a local mbarrier
__device__ __forceinline__ void tma_load_2d_sa(uint32_t smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_sa(const CUtensorMap* tmap, uint32_t smem_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(smem_addr),
               "r"(c0), "r"(c1)
               : "memory");
}

constexpr int EPI_COLS = 32;  // accumulator columns per staging step (32 fp32 = one 128-byte row segment)

// Row phase: this thread's accumulator row (32 fp32 columns) -> swizzled staging tile.  Pure transpose.
__device__ __forceinline__ void epi_stage_rows(uint32_t stg, uint32_t taddr, int lane) {
  uint32_t raw[32];
  tmem_ld_32x32b_x32(taddr, raw);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 8; ++c)
    st_shared_v4(stg + stg_off(lane, c), raw[4 * c], raw[4 * c + 1], raw[4 * c + 2], raw[4 * c + 3]);
}

// Coalesced phase: lane -> (row = it*4 + lane/8, 4 fp32 columns = chunk lane%8).  Every shared and global load of
// the step is issued before the first use (8 rows in flight per lane); the bias of the lane's 4 fixed columns is
// loaded and rounded once per step; the math of the 8 rows is independent, so the scheduler has 32 chains to overlap.
template <int EPI, int IT0, int NIT, bool FULL>
__device__ __forceinline__ void epi_write_rows(const Params& p, uint32_t stg, int row_base, int col0, int lane,
                                               float4& csum, const float4 bias4, float (&rdot)[8]) {
  const int c = lane & 7, rsub = lane >> 3;
  const int col = col0 + c * 4;
  uint4 val[8];
#pragma unroll
  for (int it = IT0; it < IT0 + NIT; ++it) val[it] = ld_shared_v4(stg + stg_off(it * 4 + rsub, c));
  if (!FULL && col >= p.N) return;
  const int row0 = row_base + rsub;  // this lane's rows are row0 + 4*it
  // FULL: the whole 32x32 step is inside the matrix, no per-row bounds checks
  auto row_ok = [&](int it) { return FULL || (row0 + it * 4 < p.M); };
  constexpr bool kSide16 = (EPI == TAE_EPI_F32_RESID || EPI == TAE_EPI_F32_ACC);
  uint4 side[8];
  if (kSide16) {
    const bool want = (EPI == TAE_EPI_F32_RESID) || (p.beta && p.splits <= 1);
    const float* sbase = (EPI == TAE_EPI_F32_RESID) ? p.resid : reinterpret_cast<const float*>(p.out);
    const size_t sld = (EPI == TAE_EPI_F32_RESID) ? (size_t)p.ldr : (size_t)p.ldo;
    const bool wrap = (EPI == TAE_EPI_F32_RESID) && p.resid_rows < p.M;  // pos-embed broadcast over the batch
#pragma unroll
    for (int it = IT0; it < IT0 + NIT; ++it) {
      side[it] = make_uint4(0u, 0u, 0u, 0u);
      if (want && row_ok(it)) {
        const int grow = row0 + it * 4;
        const int srow = wrap ? grow % p.resid_rows : grow;
        side[it] = *reinterpret_cast<const uint4*>(sbase + (size_t)srow * sld + col);
      }
    }
  } else if (EPI == TAE_EPI_BF16_DGELU || EPI == TAE_EPI_BF16_ROWDOT) {
    const bf16* abase = p.aux + (size_t)row0 * p.ldaux + col;
#pragma unroll
    for (int it = IT0; it < IT0 + NIT; ++it) {
      side[it] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok(it)) {
        const uint2 h = ld_nc_v2(abase + (size_t)(it * 4) * p.ldaux);
        side[it].x = h.x;
        side[it].y = h.y;
      }
    }
  }
  constexpr int kOutBytes = (EPI == TAE_EPI_F32_RESID || EPI == TAE_EPI_F32_ACC) ? 4 : 2;
  const size_t off0 = ((size_t)row0 * p.ldo + col) * kOutBytes;
  const size_t rstride = (size_t)p.ldo * 4 * kOutBytes;  // 4 rows
  char* obase = reinterpret_cast<char*>(p.out) + off0;
  char* obase2 = reinterpret_cast<char*>(p.out2) + off0;
#pragma unroll
  for (int it = IT0; it < IT0 + NIT; ++it) {
    if (!row_ok(it)) continue;
    const float a0 = __uint_as_float(val[it].x), a1 = __uint_as_float(val[it].y);
    const float a2 = __uint_as_float(val[it].z), a3 = __uint_as_float(val[it].w);
    char* optr = obase + it * rstride;
    if (EPI == TAE_EPI_BF16) {
      *reinterpret_cast<uint2*>(optr) =
          make_uint2(pack_bf16x2(a0 + bias4.x, a1 + bias4.y), pack_bf16x2(a2 + bias4.z, a3 + bias4.w));
    } else if (EPI == TAE_EPI_BF16_ROWDOT) {
      const uint32_t o01 = pack_bf16x2(a0 + bias4.x, a1 + bias4.y), o23 = pack_bf16x2(a2 + bias4.z, a3 + bias4.w);
      *reinterpret_cast<uint2*>(optr) = make_uint2(o01, o23);
      // dot product of the ROUNDED outputs with aux over this lane's 4 columns (the attention kernel sees those values)
      const float2 r01 = unpack_bf16x2(o01), r23 = unpack_bf16x2(o23);
      const float2 m01 = unpack_bf16x2(side[it].x), m23 = unpack_bf16x2(side[it].y);
      rdot[it] += (r01.x * m01.x + r01.y * m01.y) + (r23.x * m23.x + r23.y * m23.y);
    } else if (EPI == TAE_EPI_BF16_GELU) {
      // h = bf16(acc + bias) is the reference's fc1 output; GELU and its derivative are evaluated on that rounded value
#ifndef TAE_GELU_SCALAR
      // packed-pair evaluation (FFMA2/FMUL2): same values as the scalar path, 2/3 of its issue slots
      uint32_t g01, gp01, g23, gp23;
      gelu_and_grad_pair(f2_pack(a0, a1), f2_pack(bias4.x, bias4.y), g01, gp01);
      gelu_and_grad_pair(f2_pack(a2, a3), f2_pack(bias4.z, bias4.w), g23, gp23);
      if (p.out != nullptr) *reinterpret_cast<uint2*>(optr) = make_uint2(gp01, gp23);  // NULL: inference, gelu' not kept
      *reinterpret_cast<uint2*>(obase2 + it * rstride) = make_uint2(g01, g23);
#else
      float g[4], gp[4];
      const float2 h01 = round_bf16x2(a0 + bias4.x, a1 + bias4.y), h23 = round_bf16x2(a2 + bias4.z, a3 + bias4.w);
      gelu_and_grad_fast(h01.x, g[0], gp[0]);
      gelu_and_grad_fast(h01.y, g[1], gp[1]);
      gelu_and_grad_fast(h23.x, g[2], gp[2]);
      gelu_and_grad_fast(h23.y, g[3], gp[3]);
      if (p.out != nullptr) *reinterpret_cast<uint2*>(optr) = make_uint2(pack_bf16x2(gp[0], gp[1]), pack_bf16x2(gp[2], gp[3]));
      *reinterpret_cast<uint2*>(obase2 + it * rstride) = make_uint2(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]));
#endif
    } else if (EPI == TAE_EPI_BF16_DGELU) {
      const float2 m01 = unpack_bf16x2(side[it].x), m23 = unpack_bf16x2(side[it].y);
      // bf16(acc) first: the dgrad GEMM's own output rounding in the reference; aux holds gelu'(h) from the forward
      const float2 r01 = round_bf16x2(a0, a1), r23 = round_bf16x2(a2, a3);
      const uint32_t o01 = pack_bf16x2(r01.x * m01.x, r01.y * m01.y);
      const uint32_t o23 = pack_bf16x2(r23.x * m23.x, r23.y * m23.y);
      *reinterpret_cast<uint2*>(optr) = make_uint2(o01, o23);
      // column sums of the ROUNDED output (the bias gradient the reference derives from its bf16 dh)
      const float2 s01 = unpack_bf16x2(o01), s23 = unpack_bf16x2(o23);
      csum.x += s01.x;
      csum.y += s01.y;
      csum.z += s23.x;
      csum.w += s23.y;
    } else if (EPI == TAE_EPI_F32_RESID) {
      const float2 r01 = round_bf16x2(a0 + bias4.x, a1 + bias4.y), r23 = round_bf16x2(a2 + bias4.z, a3 + bias4.w);
      float4 o;
      o.x = __uint_as_float(side[it].x) + r01.x;
      o.y = __uint_as_float(side[it].y) + r01.y;
      o.z = __uint_as_float(side[it].z) + r23.x;
      o.w = __uint_as_float(side[it].w) + r23.y;
      *reinterpret_cast<float4*>(optr) = o;
    } else {  // TAE_EPI_F32_ACC
      float* dst = reinterpret_cast<float*>(optr);
      if (p.splits > 1) {
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(a0, a1, a2, a3));
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(a0 + __uint_as_float(side[it].x), a1 + __uint_as_float(side[it].y),
                                                      a2 + __uint_as_float(side[it].z), a3 + __uint_as_float(side[it].w));
      }
    }
  }
}

template <int EPI, int IT0 = 0, int NIT = 8>
__device__ __forceinline__ void epi_write_coalesced(const Params& p, uint32_t stg, int row_base, int col0, int lane,
                                                    float4& csum, const float4 bias4, float (&rdot)[8]) {
  if (row_base + 32 <= p.M && col0 + 32 <= p.N)  // warp-uniform
    epi_write_rows<EPI, IT0, NIT, true>(p, stg, row_base, col0, lane, csum, bias4, rdot);
  else
    epi_write_rows<EPI, IT0, NIT, false>(p, stg, row_base, col0, lane, csum, bias4, rdot);
}

// Row-dot by-product (TAE_EPI_BF16_ROWDOT): after the second 32-column step of a 64-column head, fold the 8 lanes that
// share a row (3 shuffles) and let lane c == 0 write rowdot[image, head, token] for its 8 rows.
__device__ __forceinline__ void epi_flush_rowdot(const Params& p, float (&rdot)[8], int row_base, int col0, int lane) {
  // 8 partial sums (one per row iteration) on each of the 8 lanes c = lane & 7 that share rows: a halving butterfly
  // (4 + 2 + 1 shuffles) leaves lane c with the complete sum of row iteration it = c.
  const int c = lane & 7;
  float a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // exchange across bit 2 of c: keep its[0..3] if bit clear, its[4..7] if set
    const float keep = (c & 4) ? rdot[4 + i] : rdot[i];
    const float send = (c & 4) ? rdot[i] : rdot[4 + i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {  // bit 1
    const float keep = (c & 2) ? a[2 + i] : a[i];
    const float send = (c & 2) ? a[i] : a[2 + i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float keep = (c & 1) ? b[1] : b[0];
  const float send = (c & 1) ? b[0] : b[1];
  const float v = keep + __shfl_xor_sync(0xffffffffu, send, 1);  // row iteration it == c
  const int grow = row_base + c * 4 + (lane >> 3);
  if (grow < p.M) {
    const int img = grow / p.rd_tokens, tok = grow - img * p.rd_tokens;
    p.rowdot[((size_t)img * (p.N >> 6) + (col0 >> 6)) * p.rd_tokens + tok] = v;
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) rdot[it] = 0.f;
}

// The lane's 4 bias columns of a step, rounded to bf16 (autocast hands the GEMM a bf16 copy of the fp32 bias).
// Issued BEFORE the TMEM load of the step so that its latency hides behind it.
template <int EPI>
__device__ __forceinline__ float4 epi_load_bias(const Params& p, int col0, int lane) {
  const int col = col0 + (lane & 7) * 4;
  if (EPI == TAE_EPI_F32_ACC || EPI == TAE_EPI_BF16_DGELU || p.bias == nullptr || col >= p.N)
    return make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  const float2 lo = round_bf16x2(b.x, b.y), hi = round_bf16x2(b.z, b.w);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Column-sum by-product: lanes {c, c+8, c+16, c+24} hold the partial sums of the same 4 columns over the 32 rows of
// the step; fold them with two shuffles and let lanes 0-7 write one 128-byte row of the partial-sum matrix.
__device__ __forceinline__ void epi_flush_colsum(const Params& p, float4 csum, int row_base, int col0, int lane) {
  csum.x += __shfl_xor_sync(0xffffffffu, csum.x, 8);
  csum.y += __shfl_xor_sync(0xffffffffu, csum.y, 8);
  csum.z += __shfl_xor_sync(0xffffffffu, csum.z, 8);
  csum.w += __shfl_xor_sync(0xffffffffu, csum.w, 8);
  csum.x += __shfl_xor_sync(0xffffffffu, csum.x, 16);
  csum.y += __shfl_xor_sync(0xffffffffu, csum.y, 16);
  csum.z += __shfl_xor_sync(0xffffffffu, csum.z, 16);
  csum.w += __shfl_xor_sync(0xffffffffu, csum.w, 16);
  const int col = col0 + (lane & 7) * 4;
  if (lane < 8 && col < p.N && row_base < p.M)
    *reinterpret_cast<float4*>(p.colsum_part + (size_t)(row_base >> 5) * p.N + col) = csum;
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
struct WorkItem {
  int mt, nt, kb0, kb1, split;
};
__device__ __forceinline__ WorkItem decode_work(const Params& p, int w) {
  WorkItem it;
  it.nt = w % p.n_tiles;
  const int t = w / p.n_tiles;
  it.mt = t % p.m_tiles;
  it.split = t / p.m_tiles;
  it.kb0 = it.split * p.kb_per_split;
  it.kb1 = min(p.kb_total, it.kb0 + p.kb_per_split);
  return it;
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * STAGES;     // [NUM_ACC]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + NUM_ACC;  // [NUM_ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NUM_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_work = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        const int m0 = it.mt * BLOCK_M, n0 = it.nt * BLOCK_N;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d(sa + j * MN_BOX_BYTES, &tmap_a, &full_bar[stage], m0 + j * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * MN_BOX_BYTES, &tmap_b, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = p.a_mn ? desc_mn_major(a_base, k) : desc_k_major(a_base, k);
            const uint64_t bdesc = p.b_mn ? desc_mn_major(b_base, k) : desc_k_major(b_base, k);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == NUM_ACC) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int ew = warp - 4;
    const int q = ew & 3;       // TMEM lane quarter this warp may access (== warp % 4)
    const int half = ew >> 2;   // which 128-column half of the accumulator
    const uint32_t stg = smem_u32(smem + STAGES * STAGE_BYTES + SMEM_BARRIER_BYTES + ew * STG_BYTES_PER_WARP);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int row_base = it.mt * BLOCK_M + q * 32;
      constexpr int COLS = EPI_COLS;
      float rdot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 128 / COLS; ++c) {
        const int col0 = it.nt * BLOCK_N + half * 128 + c * COLS;
        if (col0 >= p.N) break;  // warp-uniform
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + half * 128 + c * COLS);
        const float4 bias4 = epi_load_bias<EPI>(p, col0, lane);
        epi_stage_rows(stg, taddr, lane);
        __syncwarp();
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        epi_write_coalesced<EPI>(p, stg, row_base, col0, lane, csum, bias4, rdot);
        if (EPI == TAE_EPI_BF16_DGELU && p.colsum_part != nullptr) epi_flush_colsum(p, csum, row_base, col0, lane);
        if (EPI == TAE_EPI_BF16_ROWDOT && (c & 1)) epi_flush_rowdot(p, rdot, row_base, col0, lane);
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == NUM_ACC) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant: a 2-CTA cluster (two SMs of one TPC) computes a 256x256 tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 n-rows) — the tensor cores of
// both SMs read both halves — so shared-memory traffic per SM drops by a third and a stage is 32 KB.
// The leader (even) CTA issues every MMA; both CTAs run a TMA producer (its bytes are credited to the leader's
// `full` barrier) and the epilogue for their own 128 accumulator rows.
// ---------------------------------------------------------------------------------------------
constexpr int B2_STAGE_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;          // 16 KB: this CTA's half of the B tile
constexpr int STAGE2_BYTES = A_STAGE_BYTES + B2_STAGE_BYTES;         // 32 KB

// Two shapes of the kernel: EW = 8 epilogue warps (128 columns each, 32 KB of transpose staging, 6 pipeline stages) for
// main-loop-bound GEMMs and the row-layout GELU epilogue, EW = 16 (64 columns each, 64 KB staging, 5 stages) when the
// generic epilogue carries the work (GELU', row-dot and residual epilogues of short-K GEMMs).
template <int EW>
struct Cfg2 {
  static constexpr int kStages = EW == 8 ? 6 : 5;
  static constexpr int kThreads = 128 + EW * 32;
  static constexpr int kStagingBytes = EW * 32 * 128;
  static constexpr int kBiasBytes =  // bf16[32] per epilogue warp (row-layout epilogues)
      ((TAE_GELU_TMA_EPI && EW == TAE_GELU_EW) || EW == 8) ? EW * 64 : 0;
  static constexpr int kSmemBytes = kStages * STAGE2_BYTES + SMEM_BARRIER_BYTES + kStagingBytes + kBiasBytes + 1024;
  static constexpr int kColsPerWarp = 256 / (EW / 4);
};

template <int EPI, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg2<EW>::kThreads, 1)
gemm_bf16_tcgen05_2sm(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o2,
                      const Params p) {
  constexpr int STAGES2 = Cfg2<EW>::kStages;
  constexpr bool kGeluTma = TAE_GELU_TMA_EPI && EPI == TAE_EPI_BF16_GELU && EW == TAE_GELU_EW;
  constexpr bool kRowTma = kGeluTma;
  constexpr bool kDgeluTma = TAE_DGELU_TMA_EPI && EPI == TAE_EPI_BF16_DGELU && EW == 8;
  constexpr bool kRowdotTma = EPI == TAE_EPI_BF16_ROWDOT && EW == 8;
  constexpr bool kAuxTma = kDgeluTma || kRowdotTma;  // epilogues whose bf16 aux operand arrives by TMA load
  constexpr bool kInTma = kAuxTma;
  constexpr int NUM_EPI_WARPS2 = EW;
  constexpr int COLS_PER_WARP = Cfg2<EW>::kColsPerWarp;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);
  uint64_t* full_bar = bars;                                  // [STAGES2]  (the leader's copy is the live one)
  uint64_t* empty_bar = bars + STAGES2;                       // [STAGES2]  (both CTAs, multicast commit)
  uint64_t* tmem_full_bar = bars + 2 * STAGES2;               // [NUM_ACC]  (both CTAs, multicast commit)
  uint64_t* tmem_empty_bar = bars + 2 * STAGES2 + NUM_ACC;    // [NUM_ACC]  (leader's copy, 16 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 2 * NUM_ACC);
  // Dynamic tile scheduler.  A persistent kernel with a STATIC tile list doubles its run time as soon as one cluster
  // cannot become resident (e.g. NCCL's all-reduce CTAs hold a few SMs during the DDP backward): that cluster starts
  // when the others have finished and then works through its whole share alone.  Here the leader's producer thread
  // draws work items from a global counter and publishes them to both CTAs through a 16-deep ring (index + mbarrier
  // per slot, filled at cluster scope); a cluster that starts late simply finds the counter exhausted.
  constexpr int SCHED_RING = 16;  // > the producer's maximum lead over the epilogue (6 smem stages + 2 accumulators)
  uint64_t* sched_full = bars + 2 * STAGES2 + 2 * NUM_ACC + 1;              // [SCHED_RING]
  volatile int* sched_ring = reinterpret_cast<volatile int*>(sched_full + SCHED_RING);  // [SCHED_RING]
  // row-layout GELU' epilogue: one "aux tile has landed" barrier per staging box, two boxes per epilogue warp
  uint64_t* aux_bar = sched_full + SCHED_RING + SCHED_RING * sizeof(int) / sizeof(uint64_t);    // [2 * EW]
  static_assert((2 * STAGES2 + 2 * NUM_ACC + 1 + SCHED_RING + SCHED_RING / 2 + (kInTma ? 2 * EW : 0)) * 8 <= SMEM_BARRIER_BYTES,
                "barrier region overflow");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_work = p.m_tiles * p.n_tiles * p.splits;  // m_tiles counts 256-row tiles here
  const bool dynamic = p.sched != nullptr;
  // i-th work item of this cluster, as seen by a CONSUMER role (MMA issuer, epilogue warps, the peer's producer)
  auto work_at = [&](int i) -> int {
    if (!dynamic) return cluster_id + i * num_clusters;
    mbar_wait_acquire_cluster(&sched_full[i & (SCHED_RING - 1)], (uint32_t)((i / SCHED_RING) & 1));
    return sched_ring[i & (SCHED_RING - 1)];
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * NUM_EPI_WARPS2);
    }
    for (int r = 0; r < SCHED_RING; ++r) mbar_init(&sched_full[r], 1);
    if constexpr (kInTma) {
      for (int r = 0; r < 2 * EW; ++r) mbar_init(&aux_bar[r], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tcgen05_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote signal can arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // the leader draws work items ONE AHEAD: the atomic's round trip to L2 overlaps the loads of the current tile
      int w_ahead = (dynamic && rank == 0) ? atomicAdd(p.sched, 1) : 0;
      for (int i = 0;; ++i) {
        int w;
        if (dynamic && rank == 0) {
          // publish the item to both CTAs of the pair, then request the next one
          w = w_ahead;
          if (w < total_work) w_ahead = atomicAdd(p.sched, 1);
          const int slot = i & (SCHED_RING - 1);
          sched_ring[slot] = w;
          st_shared_cluster_u32(const_cast<int*>(&sched_ring[slot]), 1, (uint32_t)w);
          mbar_arrive_release_cluster(&sched_full[slot], 0);
          mbar_arrive_release_cluster(&sched_full[slot], 1);
        } else {
          w = work_at(i);
        }
        if (w >= total_work) break;
        const WorkItem it = decode_work(p, w);
        const int m0 = it.mt * (2 * BLOCK_M) + (int)rank * BLOCK_M;
        const int n0 = it.nt * p.tile_n + (int)rank * (p.tile_n / 2);
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * STAGE2_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          // bytes of both CTAs: 128 rows of A and tile_n / 2 rows of B each (a ragged box still delivers all its bytes)
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (uint32_t)(A_STAGE_BYTES + (p.tile_n / 2) * BLOCK_K * 2));
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d_2sm(sa + j * MN_BOX_BYTES, &tmap_a, &full_bar[stage], m0 + j * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 128; ++j)
              tma_load_2d_2sm(sb + j * MN_BOX_BYTES, &tmap_b, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (dynamic && rank == 0) {
        // every cluster has drawn its last (out-of-range) item before it counts itself done: the last one re-arms the
        // counters for the next launch that uses this slot
        if (atomicAdd(p.sched + 1, 1) == num_clusters - 1) {
          p.sched[0] = 0;
          p.sched[1] = 0;
          __threadfence();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, p.tile_n, p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int i = 0;; ++i) {
        const int w = work_at(i);
        if (w >= total_work) break;
        const WorkItem it = decode_work(p, w);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * STAGE2_BYTES);
          const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = p.a_mn ? desc_mn_major(a_base, k) : desc_k_major(a_base, k);
            const uint64_t bdesc = p.b_mn ? desc_mn_major(b_base, k) : desc_k_major(b_base, k);
            umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2sm(&tmem_full_bar[acc]);
        if (++acc == NUM_ACC) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int ew = warp - 4;
    const int q = ew & 3;       // TMEM lane quarter (== warp % 4)
    const int cg = ew >> 2;     // which column group of the accumulator
    const uint32_t stg = smem_u32(smem + STAGES2 * STAGE2_BYTES + SMEM_BARRIER_BYTES + ew * STG_BYTES_PER_WARP);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t aux_phase = 0;  // kAuxTma: parity of the two aux-box barriers (bit b = box b)
    for (int i = 0;; ++i) {
      const int w = work_at(i);
      if (w >= total_work) break;
      const WorkItem it = decode_work(p, w);
      const int row_base = it.mt * (2 * BLOCK_M) + (int)rank * BLOCK_M + q * 32;
      const int n_lim = min(p.N, (it.nt + 1) * p.tile_n);  // first column this tile does not own (tile_n % 32 == 0)
      if constexpr (kInTma) {
        // the first input box of the tile (gelu'(h), the row-dot operand, or the fp32 residual: 2 KB each) is requested
        // before the wait for the accumulator
        const int colw = it.nt * p.tile_n + cg * COLS_PER_WARP;
        if (colw < n_lim && row_base < p.M && elect_one()) {
          tma_store_wait_read();  // this warp's earlier TMA stores have finished reading both boxes
          mbar_expect_tx(&aux_bar[ew * 2], 2048u);
          tma_load_2d_sa(stg, &tmap_o2, &aux_bar[ew * 2], colw, row_base);
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      if constexpr (kAuxTma) {
        // Row-layout GELU' epilogue: out = bf16(bf16(acc) * gelu'(h)).  Step c works in staging box c & 1: the aux
        // tile was TMA-loaded into it one step earlier, every thread multiplies its own row in place, the box leaves by
        // TMA store, and the per-32-row column sums of the rounded products (the fc1 bias gradient) are read back from
        // the staged box: lane -> (column pair, odd/even rows), one shuffle to fold the two row halves.
        constexpr int NSTEP = COLS_PER_WARP / EPI_COLS;
        const int colw = it.nt * p.tile_n + cg * COLS_PER_WARP;
        const bool active = colw < n_lim && row_base < p.M;  // warp-uniform
        bool released = false;
        // Row-dot variant: out = bf16(acc + bias) written over the aux tile, and dot = sum over the 64 columns of a head
        // of out * aux stays in this thread (it owns the row); written after the head's second step.
        [[maybe_unused]] const uint32_t bias_sa =
            smem_u32(smem + STAGES2 * STAGE2_BYTES + SMEM_BARRIER_BYTES + Cfg2<EW>::kStagingBytes) + (uint32_t)ew * 64u;
        [[maybe_unused]] float dot = 0.f;
#pragma unroll 1
        for (int c = 0; c < NSTEP; ++c) {
          const int col0 = colw + c * EPI_COLS;
          if (!active || col0 >= n_lim) break;
          const bool last = (c == NSTEP - 1) || (col0 + EPI_COLS >= n_lim);
          const int b = c & 1;
          const uint32_t box = stg + (uint32_t)b * 2048u;
          const uint32_t taddr =
              tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + cg * COLS_PER_WARP + c * EPI_COLS);
          [[maybe_unused]] float bv = 0.f;
          if constexpr (kRowdotTma) {
            if (p.bias != nullptr && col0 + lane < p.N) bv = __ldg(p.bias + col0 + lane);
          }
          uint32_t raw[32];
          tmem_ld_32x32b_x32(taddr, raw);
          if constexpr (kRowdotTma) st_shared_u16(bias_sa + (uint32_t)lane * 2u, __bfloat16_as_ushort(__float2bfloat16_rn(bv)));
          if (!last && elect_one()) {  // next step's aux tile into the other box (its last store has been read)
            tma_store_wait_read();
            mbar_expect_tx(&aux_bar[ew * 2 + (b ^ 1)], 2048u);
            tma_load_2d_sa(stg + (uint32_t)(b ^ 1) * 2048u, &tmap_o2, &aux_bar[ew * 2 + (b ^ 1)], col0 + EPI_COLS, row_base);
          }
          tmem_ld_wait();
          if (last) tcgen05_fence_before();
          __syncwarp();
          if (last) {
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            released = true;
          }
          mbar_wait(&aux_bar[ew * 2 + b], (aux_phase >> b) & 1u);
          aux_phase ^= 1u << b;
          [[maybe_unused]] uint4 bq[4];
          if constexpr (kRowdotTma) {
#pragma unroll
            for (int k = 0; k < 4; ++k) bq[k] = ld_shared_v4(bias_sa + (uint32_t)k * 16u);  // after the __syncwarp above
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint4 m = ld_shared_v4(box + stg64_off(lane, k));
            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = 8 * k + 2 * j;
              if constexpr (kRowdotTma) {
                const uint32_t bw = (&bq[k].x)[j];
                float s0, s1, d0, d1;
                f2_unpack(f2_add(f2_pack(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1])),
                                 f2_pack(__uint_as_float(bw << 16), __uint_as_float(bw & 0xffff0000u))), s0, s1);
                ow[j] = pack_bf16x2(s0, s1);
                // dot product of the ROUNDED outputs with aux (the attention kernel sees those values)
                f2_unpack(f2_mul(f2_pack(__uint_as_float(ow[j] << 16), __uint_as_float(ow[j] & 0xffff0000u)),
                                 f2_pack(__uint_as_float(mw[j] << 16), __uint_as_float(mw[j] & 0xffff0000u))), d0, d1);
                dot += d0 + d1;
                continue;
              }
              // bf16(acc) first: the dgrad GEMM's own output rounding in the reference
              const float2 r = round_bf16x2(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1]));
              float o0, o1;
              f2_unpack(f2_mul(f2_pack(r.x, r.y), f2_pack(__uint_as_float(mw[j] << 16), __uint_as_float(mw[j] & 0xffff0000u))),
                        o0, o1);
              ow[j] = pack_bf16x2(o0, o1);
            }
            st_shared_v4(box + stg64_off(lane, k), ow[0], ow[1], ow[2], ow[3]);
          }
          if constexpr (kRowdotTma) {
            if (c & 1) {  // second half of a 64-column head: rowdot[image, head, token] of this thread's row
              const int grow = row_base + lane;
              if (grow < p.M) {
                const int img = grow / p.rd_tokens, tok = grow - img * p.rd_tokens;
                p.rowdot[((size_t)img * (p.N >> 6) + (col0 >> 6)) * p.rd_tokens + tok] = dot;
              }
              dot = 0.f;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_2d_sa(&tmap_o, box, col0, row_base);
            tma_store_commit();
          }
          if (p.colsum_part != nullptr) {
            const int jp = lane & 15, rh = lane >> 4;  // columns 2 jp, 2 jp + 1; rows rh, rh + 2, ...
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
              const int r = 2 * rr + rh;
              const uint32_t wv = ld_shared_u32(box + stg64_off(r, jp >> 2) + (uint32_t)(jp & 3) * 4u);
              s0 += __uint_as_float(wv << 16);
              s1 += __uint_as_float(wv & 0xffff0000u);
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            const int col = col0 + 2 * jp;
            if (lane < 16 && col < p.N)
              *reinterpret_cast<float2*>(p.colsum_part + (size_t)(row_base >> 5) * p.N + col) = make_float2(s0, s1);
          }
          __syncwarp();  // every lane is done with the box before a later TMA load may overwrite it
        }
        if (!released) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        }
        if (++acc == NUM_ACC) {
          acc = 0;
          acc_phase ^= 1u;
        }
        continue;
      }
      if constexpr (kRowTma) {
        // Row-layout epilogue: thread = accumulator row, 32 consecutive columns per step.  gelu(h) and gelu'(h) are packed
        // to bf16 straight from the tcgen05.ld registers into two [32 rows x 32 cols] staging boxes (64 B per row,
        // 16-byte chunks XOR-swizzled) and leave through two TMA stores per step, which also clip the M / N edges.
        // The accumulator buffer is handed back to the MMA issuer as soon as the tile's last tcgen05.ld has landed.
        const uint32_t bias_sa = smem_u32(smem + STAGES2 * STAGE2_BYTES + SMEM_BARRIER_BYTES + Cfg2<EW>::kStagingBytes) +
                                 (uint32_t)ew * 64u;
        constexpr int NSTEP = COLS_PER_WARP / EPI_COLS;
        bool released = false;
#pragma unroll 1
        for (int c = 0; c < NSTEP; ++c) {
          const int col0 = it.nt * p.tile_n + cg * COLS_PER_WARP + c * EPI_COLS;
          if (col0 >= n_lim) break;  // warp-uniform
          const bool last = (c == NSTEP - 1) || (col0 + EPI_COLS >= n_lim);
          const uint32_t taddr =
              tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + cg * COLS_PER_WARP + c * EPI_COLS);
          // the step's 32 bias values are warp-uniform: lane j fetches and rounds column col0 + j, the warp shares them
          // through 64 bytes of shared memory (autocast hands the GEMM a bf16 copy of the fp32 bias)
          float bv = 0.f;
          if (p.bias != nullptr && col0 + lane < p.N) bv = __ldg(p.bias + col0 + lane);
          uint32_t raw[32];
          tmem_ld_32x32b_x32(taddr, raw);
          st_shared_u16(bias_sa + (uint32_t)lane * 2u, __bfloat16_as_ushort(__float2bfloat16_rn(bv)));
          tmem_ld_wait();
          if (last) tcgen05_fence_before();
          // the staging boxes about to be overwritten are no longer being read (GELU rewrites both boxes every step)
          if (elect_one()) tma_store_wait_read();
          __syncwarp();
          if (last) {
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            released = true;
          }
          uint4 bq[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) bq[k] = ld_shared_v4(bias_sa + (uint32_t)k * 16u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 8 columns = one 16-byte chunk of each output row
            const uint32_t bw[4] = {bq[k].x, bq[k].y, bq[k].z, bq[k].w};
            uint32_t gw[4], pw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = 8 * k + 2 * j;
              gelu_and_grad_pair(f2_pack(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1])),
                                 f2_pack(__uint_as_float(bw[j] << 16), __uint_as_float(bw[j] & 0xffff0000u)), gw[j], pw[j]);
            }
            st_shared_v4(stg + stg64_off(lane, k), gw[0], gw[1], gw[2], gw[3]);          // gelu(h)  -> out2
            if (p.out != nullptr)  // NULL: inference (forward_encoder under no_grad): gelu'(h) is not kept
              st_shared_v4(stg + 2048u + stg64_off(lane, k), pw[0], pw[1], pw[2], pw[3]);  // gelu'(h) -> out
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (row_base < p.M && elect_one()) {
            tma_store_2d_sa(&tmap_o2, stg, col0, row_base);
            if (p.out != nullptr) tma_store_2d_sa(&tmap_o, stg + 2048u, col0, row_base);
            tma_store_commit();
          }
        }
        if (!released) {  // this warp's columns lie outside the matrix
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        }
        if (++acc == NUM_ACC) {
          acc = 0;
          acc_phase ^= 1u;
        }
        continue;
      }
      float rdot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP / EPI_COLS; ++c) {
        const int col0 = it.nt * p.tile_n + cg * COLS_PER_WARP + c * EPI_COLS;
        if (col0 >= n_lim) break;  // warp-uniform
        const uint32_t taddr =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + cg * COLS_PER_WARP + c * EPI_COLS);
        const float4 bias4 = epi_load_bias<EPI>(p, col0, lane);
        epi_stage_rows(stg, taddr, lane);
        __syncwarp();
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EW == 8) {
          epi_write_coalesced<EPI, 0, 8>(p, stg, row_base, col0, lane, csum, bias4, rdot);
        } else {  // two passes of 16 rows keep the live register set under the 640-thread budget
          epi_write_coalesced<EPI, 0, 4>(p, stg, row_base, col0, lane, csum, bias4, rdot);
          epi_write_coalesced<EPI, 4, 4>(p, stg, row_base, col0, lane, csum, bias4, rdot);
        }
        if (EPI == TAE_EPI_BF16_DGELU && p.colsum_part != nullptr) epi_flush_colsum(p, csum, row_base, col0, lane);
        if (EPI == TAE_EPI_BF16_ROWDOT && (c & 1)) epi_flush_rowdot(p, rdot, row_base, col0, lane);
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);  // the leader's MMA thread owns the accumulators
      if (++acc == NUM_ACC) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if constexpr (kRowTma || kInTma) {
      if (elect_one()) tma_store_wait_all();  // shared memory must outlive the last TMA store's reads
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
template <int EPI>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int grid, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(gemm_bf16_tcgen05<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%d) failed: %s", SMEM_BYTES, cudaGetErrorString(attr_err));
    return TAE_ERR_CUDA;
  }
  gemm_bf16_tcgen05<EPI><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, p);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

template <int EPI, int EW>
static int launch_2sm_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                          const Params& p, int clusters, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(gemm_bf16_tcgen05_2sm<EPI, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg2<EW>::kSmemBytes);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%d) failed: %s", Cfg2<EW>::kSmemBytes, cudaGetErrorString(attr_err));
    return TAE_ERR_CUDA;
  }
  gemm_bf16_tcgen05_2sm<EPI, EW><<<2 * clusters, Cfg2<EW>::kThreads, Cfg2<EW>::kSmemBytes, stream>>>(ta, tb, to, to2, p);
  TAE_CHECK_LAUNCH();
  return TAE_OK;
}

template <int EPI>
static int launch_2sm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                      const Params& p, int clusters, cudaStream_t stream) {
  // 16 epilogue warps where the generic (transposing) epilogue carries the work: the residual epilogue (and the GELU /
  // GELU' epilogues if built without their row-layout paths) when the main loop per tile is short (K <= 2048)
  constexpr bool kMayBeHeavy = (EPI == TAE_EPI_BF16_GELU && TAE_GELU_EW == 16) || EPI == TAE_EPI_F32_RESID ||
                               (EPI == TAE_EPI_BF16_DGELU && !TAE_DGELU_TMA_EPI);
  if constexpr (kMayBeHeavy) {
    if (EPI == TAE_EPI_BF16_GELU || p.K <= 2048) return launch_2sm_cfg<EPI, 16>(ta, tb, to, to2, p, clusters, stream);
  }
  return launch_2sm_cfg<EPI, 8>(ta, tb, to, to2, p, clusters, stream);
}

}  // namespace gemm
}  // namespace tae

extern "C" int tae_gemm(const tae_gemm_args* a, void* stream_) {
  using namespace tae;
  using namespace tae::gemm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  TAE_CHECK_SHAPE(a != nullptr, "tae_gemm: args is NULL");
  TAE_CHECK_SHAPE(a->M > 0 && a->N > 0 && a->K > 0, "tae_gemm: M,N,K must be positive (got %d,%d,%d)", a->M, a->N, a->K);
  TAE_CHECK_SHAPE(a->N % 8 == 0 && a->K % 8 == 0, "tae_gemm: N and K must be multiples of 8 (got N=%d K=%d)", a->N, a->K);
  TAE_CHECK_SHAPE(a->lda % 8 == 0 && a->ldb % 8 == 0, "tae_gemm: lda/ldb must be multiples of 8");
  TAE_CHECK_SHAPE(!a->a_mn_major || a->M % 8 == 0, "tae_gemm: MN-major A requires M %% 8 == 0");
  TAE_CHECK_SHAPE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                  "tae_gemm: A, B and out must be 16-byte aligned");
  TAE_CHECK_SHAPE((a->out != nullptr || (a->epilogue == TAE_EPI_BF16_GELU && a->out2 != nullptr)) && a->A != nullptr &&
                      a->B != nullptr,
                  "tae_gemm: NULL operand (out may be NULL only for TAE_EPI_BF16_GELU: inference, gelu' not kept)");
  TAE_CHECK_SHAPE(a->epilogue >= TAE_EPI_BF16 && a->epilogue <= TAE_EPI_BF16_ROWDOT, "tae_gemm: bad epilogue %d", a->epilogue);
  const bool out_f32 = (a->epilogue == TAE_EPI_F32_RESID || a->epilogue == TAE_EPI_F32_ACC);
  TAE_CHECK_SHAPE(a->ldo % (out_f32 ? 4 : 8) == 0 && a->ldo >= a->N, "tae_gemm: bad ldo %d", a->ldo);
  if (a->epilogue == TAE_EPI_BF16_GELU)
    TAE_CHECK_SHAPE(a->out2 != nullptr && (reinterpret_cast<uintptr_t>(a->out2) & 15) == 0, "tae_gemm: GELU epilogue needs out2");
  if (a->epilogue == TAE_EPI_F32_RESID)
    TAE_CHECK_SHAPE(a->resid != nullptr && a->resid_rows > 0 && a->ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(a->resid) & 15) == 0,
                    "tae_gemm: RESID epilogue needs resid/resid_rows/ldr");
  if (a->epilogue == TAE_EPI_BF16_ROWDOT)
    TAE_CHECK_SHAPE(a->rowdot != nullptr && a->rowdot_tokens > 0 && a->M % a->rowdot_tokens == 0 && a->N % 64 == 0,
                    "tae_gemm: ROWDOT epilogue needs rowdot, M %% rowdot_tokens == 0 and N %% 64 == 0");
  if (a->epilogue == TAE_EPI_BF16_DGELU || a->epilogue == TAE_EPI_BF16_ROWDOT)
    TAE_CHECK_SHAPE(a->aux != nullptr && a->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0,
                    "tae_gemm: DGELU / ROWDOT epilogues need aux/ldaux");
  if (a->colsum_partials)
    TAE_CHECK_SHAPE(a->epilogue == TAE_EPI_BF16_DGELU && (reinterpret_cast<uintptr_t>(a->colsum_partials) & 15) == 0,
                    "tae_gemm: colsum_partials is only produced by the TAE_EPI_BF16_DGELU epilogue (16-byte aligned)");
  if (a->bias) TAE_CHECK_SHAPE((reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "tae_gemm: bias must be 16-byte aligned");

  Params p{};
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  const int sms = num_sms();
  if (sms <= 0) return TAE_ERR_CUDA;
  // CTA-pair kernel (256-row tiles on two SMs) whenever there are at least two 128-row tiles of work
  const bool use2 = a->M > BLOCK_M && sms >= 2;
  const int tile_m = use2 ? 2 * BLOCK_M : BLOCK_M;
  const int units = use2 ? sms / 2 : sms;  // concurrently resident work items (clusters or CTAs)
  p.m_tiles = (a->M + tile_m - 1) / tile_m;
  // Tile width.  With few token rows (patch64 / patch128: 16 / 4 row tiles) a 256-wide tiling leaves the last wave of the
  // 74 clusters mostly empty (N = 2560: 160 tiles = 2.16 waves at M = 4096, 40 tiles = 0.54 at M = 1024).  For K-major B
  // (every forward GEMM; MN-major B is tied to 64-column TMA boxes) the width is chosen from {256, 224, 192, 160, 128} to
  // maximise wave efficiency x MMA efficiency, the latter from the measured shared-memory operand limit of the pair MMA
  // (72 B/clk/SM: 4 KB of A + 16 tile_n B of B per instruction against tile_n / 2 cycles of array time).
  p.tile_n = BLOCK_N;
#ifndef TAE_FIXED_TILE_N  // A/B builds: -DTAE_FIXED_TILE_N keeps 256-wide tiles everywhere
  if (use2 && !p.b_mn && a->epilogue != TAE_EPI_F32_ACC && a->epilogue != TAE_EPI_BF16_ROWDOT) {  // (a row-dot head = 64 columns)
    double best = 0.0;
    for (int tn = BLOCK_N; tn >= 128; tn -= 32) {
      const long tiles = (long)p.m_tiles * ((a->N + tn - 1) / tn);
      const double wave = (double)tiles / (double)(((tiles + units - 1) / units) * units);
      const double smem_cycles = (4096.0 + 16.0 * tn) / 72.0, array_cycles = tn / 2.0;
      const double score = wave * (array_cycles / (smem_cycles > array_cycles ? smem_cycles : array_cycles));
      if (score > best + (tn == BLOCK_N ? 0.0 : 0.03)) {  // a narrower tile has to pay for itself
        best = score;
        p.tile_n = tn;
      }
    }
  }
#endif
  p.n_tiles = (a->N + (use2 ? p.tile_n : BLOCK_N) - 1) / (use2 ? p.tile_n : BLOCK_N);
  p.kb_total = (a->K + BLOCK_K - 1) / BLOCK_K;
  int splits = a->splits;
  if (a->epilogue != TAE_EPI_F32_ACC) {
    TAE_CHECK_SHAPE(splits <= 1, "tae_gemm: split-K only with TAE_EPI_F32_ACC");
    splits = 1;
  } else if (splits <= 0) {
    // auto: weight-gradient GEMMs have few output tiles and a huge K.  Pick the split factor whose work-item count
    // wastes the least of the last wave (items / (waves * SMs)); ties go to the smaller factor.
    const int tiles = p.m_tiles * p.n_tiles;
    splits = 1;
    if (p.kb_total >= 32) {
      double best = (double)tiles / (double)(((tiles + units - 1) / units) * units);
      const int max_s = p.kb_total / 16 < 16 ? p.kb_total / 16 : 16;
      for (int sp = 2; sp <= max_s; ++sp) {
        const long items = (long)tiles * sp;
        const double eff = (double)items / (double)(((items + units - 1) / units) * units);
        if (eff > best + 0.02) {
          best = eff;
          splits = sp;
        }
      }
    }
  }
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.splits = splits;
  p.out = a->out;
  p.ldo = a->ldo;
  p.out2 = a->out2;
  p.bias = a->bias;
  p.resid = a->resid;
  p.ldr = a->ldr;
  p.resid_rows = a->resid_rows > 0 ? a->resid_rows : 1;
  p.aux = reinterpret_cast<const bf16*>(a->aux);
  p.ldaux = a->ldaux;
  p.beta = a->beta;
  p.colsum_part = a->colsum_partials;
  p.rowdot = a->rowdot;
  p.rd_tokens = a->rowdot_tokens > 0 ? a->rowdot_tokens : 1;

  if (a->epilogue == TAE_EPI_F32_ACC && p.splits > 1 && !a->beta) {
    // split-K partial sums are accumulated with red.global.add: start from zero
    TAE_CHECK_CUDA(cudaMemset2DAsync(a->out, (size_t)a->ldo * 4, 0, (size_t)a->N * 4, (size_t)a->M, stream));
  }

  CUtensorMap ta, tb;
  int rc;
  if (!p.a_mn)
    rc = make_tmap(&ta, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BLOCK_M);
  else
    rc = make_tmap(&ta, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, BLOCK_K);
  if (rc) return rc;
  if (!p.b_mn)
    rc = make_tmap(&tb, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, use2 ? p.tile_n / 2 : BLOCK_N);
  else
    rc = make_tmap(&tb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, BLOCK_K);
  if (rc) return rc;

  const int total = p.m_tiles * p.n_tiles * p.splits;
  CUtensorMap to{}, to2{};  // output / aux maps of the row-layout epilogues (GELU, GELU', row-dot)
  if (use2 && ((TAE_GELU_TMA_EPI && a->epilogue == TAE_EPI_BF16_GELU) ||
               (TAE_DGELU_TMA_EPI && a->epilogue == TAE_EPI_BF16_DGELU) || a->epilogue == TAE_EPI_BF16_ROWDOT)) {
    const CUtensorMapSwizzle swz = TAE_GELU_TMA_SWZ64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    if (a->out != nullptr) {
      rc = make_tmap_box(&to, a->out, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldo, 32, 32, swz);
      if (rc) return rc;
    }
    if (a->epilogue == TAE_EPI_BF16_GELU) {
      rc = make_tmap_box(&to2, a->out2, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldo, 32, 32, swz);
      if (rc) return rc;
    } else if (a->epilogue == TAE_EPI_BF16_DGELU || a->epilogue == TAE_EPI_BF16_ROWDOT) {  // the second map carries aux
      rc = make_tmap_box(&to2, a->aux, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldaux, 32, 32, swz);
      if (rc) return rc;
    }
  }
  if (use2) {
    const int clusters = total < units ? total : units;
    p.sched = total > clusters ? sched_counter_slot() : nullptr;  // one item per cluster needs no scheduler
    switch (a->epilogue) {
      case TAE_EPI_BF16: return launch_2sm<TAE_EPI_BF16>(ta, tb, to, to2, p, clusters, stream);
      case TAE_EPI_BF16_GELU: return launch_2sm<TAE_EPI_BF16_GELU>(ta, tb, to, to2, p, clusters, stream);
      case TAE_EPI_F32_RESID: return launch_2sm<TAE_EPI_F32_RESID>(ta, tb, to, to2, p, clusters, stream);
      case TAE_EPI_F32_ACC: return launch_2sm<TAE_EPI_F32_ACC>(ta, tb, to, to2, p, clusters, stream);
      case TAE_EPI_BF16_DGELU: return launch_2sm<TAE_EPI_BF16_DGELU>(ta, tb, to, to2, p, clusters, stream);
      case TAE_EPI_BF16_ROWDOT: return launch_2sm<TAE_EPI_BF16_ROWDOT>(ta, tb, to, to2, p, clusters, stream);
    }
    return TAE_ERR_SHAPE;
  }
  const int grid = total < sms ? total : sms;
  switch (a->epilogue) {
    case TAE_EPI_BF16: return launch<TAE_EPI_BF16>(ta, tb, p, grid, stream);
    case TAE_EPI_BF16_GELU: return launch<TAE_EPI_BF16_GELU>(ta, tb, p, grid, stream);
    case TAE_EPI_F32_RESID: return launch<TAE_EPI_F32_RESID>(ta, tb, p, grid, stream);
    case TAE_EPI_F32_ACC: return launch<TAE_EPI_F32_ACC>(ta, tb, p, grid, stream);
    case TAE_EPI_BF16_DGELU: return launch<TAE_EPI_BF16_DGELU>(ta, tb, p, grid, stream);
    case TAE_EPI_BF16_ROWDOT: return launch<TAE_EPI_BF16_ROWDOT>(ta, tb, p, grid, stream);
  }
  return TAE_ERR_SHAPE;
}
