/*
 * tae_b200.h — C ABI of libtae_b200.so: the sm_100a kernels behind the TAE hot path.
 *
 * The reference (eminorhan/tae) has no FFI of its own: every device instruction on its hot path is
 * reached through a PyTorch library call inside tae.py / util/misc.py.  Each entry point below
 * replaces one (or a fused group) of those call sites; the call site is cited per function as
 * reference file:line (relative to the reference repo root).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions (every symbol):
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers owned by the caller;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous, never synchronise,
 *     never allocate device memory and never read results back to the host;
 *   - return value 0 on success, negative on failure:
 *       TAE_ERR_SHAPE (-1) bad shape/alignment, TAE_ERR_WORKSPACE (-2) workspace too small,
 *       TAE_ERR_CUDA (-3) CUDA launch/driver error, TAE_ERR_UNSUPPORTED (-4) unsupported arch/config;
 *     tae_last_error_string() (thread-local) describes the last failure;
 *   - re-entrant from any host thread (autograd's backward thread differs from the forward thread);
 *   - bf16 tensors are raw uint16_t storage (`tae_bf16`), fp32 tensors are float.
 */
#ifndef TAE_B200_H_
#define TAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAE_OK 0
#define TAE_ERR_SHAPE (-1)
#define TAE_ERR_WORKSPACE (-2)
#define TAE_ERR_CUDA (-3)
#define TAE_ERR_UNSUPPORTED (-4)

typedef uint16_t tae_bf16;

/* ---- library / device ------------------------------------------------------------------- */
int tae_version(void);                       /* ABI version, currently 1 */
const char* tae_build_fingerprint(void);     /* sha256 of the sources the library was built from ("unknown" if unset) */
const char* tae_last_error_string(void);     /* thread-local, never NULL */
int tae_device_check(void);                  /* 0 iff current device is sm_100 (B200) */
int tae_num_sms(void);                       /* SM count of the current device (148 on B200), <0 on error */
/* Work scheduling of the persistent kernels (GEMM tiles, LayerNorm-backward rows).  Static lists (default) are ~0.5 %
 * faster when the GPU is not shared; dynamic lists (a global counter per launch) keep a kernel from doubling its run
 * time when some of its CTAs cannot become resident, e.g. while NCCL's all-reduce kernels hold SMs during a
 * data-parallel backward — tae_b200.ddp turns them on for world_size > 1.  Returns the previous setting (-1 = unset). */
int tae_set_dynamic_scheduling(int enable);
/* number of kernel launches issued through this library since process start (all threads) */
uint64_t tae_launch_count(void);

/* ---- GEMM (tcgen05 / TMEM / TMA) ---------------------------------------------------------
 * D[M,N] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulation in tensor memory.
 * Replaces every nn.Linear on the path and their autograd dgrad/wgrad:
 *   tae.py:74 (qkv), :81 (attn proj), :101 (fc1), :104 (fc2), :237 (dict_proj), :242 (decoder_embed),
 *   :253 (decoder_pred), :50 (PatchEmbed conv as im2col GEMM).
 * Operand storage:
 *   a_mn_major == 0: A is row-major [M, K] with leading dimension lda (k contiguous)
 *   a_mn_major == 1: A is stored transposed, row-major [K, M] with leading dimension lda (m contiguous)
 *   b_mn_major == 0: B is row-major [N, K] (nn.Linear weight layout [out,in]) with leading dimension ldb
 *   b_mn_major == 1: B is stored row-major [K, N] with leading dimension ldb (n contiguous)
 * so that   forward  y = x W^T         : A=x (0),  B=W  (0)
 *           dgrad    dx = dy W         : A=dy (0), B=W  (1)   (W is [N_fwd, K_fwd] = [K, N] here)
 *           wgrad    dW = dy^T x       : A=dy (1), B=x  (1)
 * Requirements: K % 8 == 0, N % 8 == 0, lda/ldb % 8 == 0, all pointers 16-byte aligned.
 */
enum tae_gemm_epilogue {
  /* out(bf16)[m,n] = acc + bias[n]                      (bias may be NULL) */
  TAE_EPI_BF16 = 0,
  /* h = bf16(acc + bias[n]);  out2 = bf16(gelu_erf(h))  (tae.py:101-102);  out = bf16(gelu_erf'(h)) — the only
   * thing backward needs from the pre-activation, so it is saved instead of h and the fc2 dgrad epilogue
   * (TAE_EPI_BF16_DGELU) is a plain multiply.  out == NULL: inference (forward_encoder / forward under no_grad,
   * encode.py:85) — gelu'(h) is not written */
  TAE_EPI_BF16_GELU = 1,
  /* out(f32)[m,n] = resid[(m % resid_rows), n] + float(bf16(acc + bias[n]))
   * residual add tae.py:129-130; pos-embed add tae.py:229,245 (resid_rows = tokens per image) */
  TAE_EPI_F32_RESID = 2,
  /* out(f32)[m,n] = (beta ? out[m,n] : 0) + acc; with splits > 1 the partial sums are added with
   * red.global.add (out must then be pre-initialised; beta is implied)      weight gradients */
  TAE_EPI_F32_ACC = 3,
  /* out(bf16)[m,n] = bf16( float(bf16(acc)) * aux[m,n] ),  aux = gelu_erf'(h) saved by TAE_EPI_BF16_GELU:
   * GELU backward fused into the fc2 dgrad */
  TAE_EPI_BF16_DGELU = 4,
  /* out(bf16)[m,n] = acc + bias[n] (as TAE_EPI_BF16), and in the same pass
   *   rowdot[b, h, t] = sum_{n in head h} float(out[m,n]) * float(aux[m,n]),   m = b*rowdot_tokens + t, head = 64 columns:
   * the attention backward's delta = rowsum(dO * O) emitted by the GEMM that produces dO (attn.proj dgrad, tae.py:81),
   * so the attention kernel never re-reads O and dO for it.  N % 64 == 0. */
  TAE_EPI_BF16_ROWDOT = 5
};

typedef struct tae_gemm_args {
  const tae_bf16* A;
  const tae_bf16* B;
  int32_t M, N, K;
  int32_t lda, ldb;
  int32_t a_mn_major, b_mn_major;
  int32_t epilogue;        /* enum tae_gemm_epilogue */
  void* out;               /* bf16 or fp32 per epilogue, row-major [M, N], leading dim ldo */
  int32_t ldo;
  void* out2;              /* TAE_EPI_BF16_GELU: bf16 gelu(h), same layout as out (which receives gelu'(h)) */
  const float* bias;       /* [N] fp32 or NULL */
  const float* resid;      /* TAE_EPI_F32_RESID: fp32 [resid_rows, N], leading dim ldr (may alias out) */
  int32_t ldr;
  int32_t resid_rows;
  const tae_bf16* aux;     /* TAE_EPI_BF16_DGELU: bf16 multiplier [M, N], leading dim ldaux */
  int32_t ldaux;
  int32_t beta;            /* TAE_EPI_F32_ACC */
  int32_t splits;          /* split-K factor, >= 1; only TAE_EPI_F32_ACC may use > 1; 0 = auto */
  float* colsum_partials;  /* TAE_EPI_BF16_DGELU, optional: fp32 [ceil(M/32), N]; row r receives the column sums of output
                            * rows 32r..32r+31 (of the bf16-rounded values).  tae_colsum_f32 reduces it to the bias
                            * gradient without re-reading the [M, N] output */
  float* rowdot;           /* TAE_EPI_BF16_ROWDOT: fp32 [M / rowdot_tokens, N / 64, rowdot_tokens]; aux/ldaux give the multiplier */
  int32_t rowdot_tokens;   /* tokens per image (M % rowdot_tokens == 0) */
} tae_gemm_args;

int tae_gemm(const tae_gemm_args* args, void* stream);

/* ---- LayerNorm -----------------------------------------------------------------------------
 * nn.LayerNorm(D, eps) with fp32 statistics, biased variance  (tae.py:122,125,159,168; eps=1e-6 via
 * tae.py:435).  x fp32 [rows, D] -> y bf16 [rows, D]; mean/rstd fp32 [rows] saved for backward.
 * D % 128 == 0 and D <= 8192.
 */
int tae_layernorm_fwd(const float* x, const float* gamma, const float* beta, tae_bf16* y,
                      float* mean, float* rstd, int32_t rows, int32_t D, float eps, void* stream);

/* LayerNorm backward fused with the residual-gradient add and the bf16 re-cast of the stream:
 *   dx      = LN'(dy; x, mean, rstd, gamma)
 *   dres_out = dres_in + dx                 (fp32; dres_in may be NULL (= 0) and may alias dres_out)
 *   dres_out_bf16 = bf16(dres_out)          (may be NULL)
 * Column partial sums (dgamma, dbeta, colsum(bf16(dres_out))) go to `partials`
 * (fp32 [tae_layernorm_bwd_num_partials(rows, D)][3][D]); tae_layernorm_bwd_finalize reduces them into
 * dgamma/dbeta/dcolsum.  `accumulate` is a bit mask: bit 0 -> dgamma +=, bit 1 -> dbeta +=, bit 2 -> dcolsum +=
 * (a clear bit overwrites).  Any of the three outputs may be NULL.
 */
int tae_layernorm_bwd_num_partials(int32_t rows, int32_t D);
int tae_layernorm_bwd(const tae_bf16* dy, const float* x, const float* mean, const float* rstd,
                      const float* gamma, const float* dres_in, float* dres_out,
                      tae_bf16* dres_out_bf16, float* partials, int32_t rows, int32_t D, void* stream);
int tae_layernorm_bwd_finalize(const float* partials, int32_t num_partials, int32_t D,
                               float* dgamma, float* dbeta, float* dcolsum, int32_t accumulate,
                               void* stream);

/* ---- Attention (short sequences held entirely in shared memory) -----------------------------
 * F.scaled_dot_product_attention(q, k, v), non-causal, no mask, scale 1/sqrt(hd)   (tae.py:78),
 * including the qkv un-bind permute (tae.py:74-75) and the head merge (tae.py:80):
 *   qkv  bf16 [B*N, 3*H*hd]  (row = b*N + n; col = which*H*hd + h*hd + d; which 0=q,1=k,2=v)
 *   out  bf16 [B*N, H*hd]    (col = h*hd + d)
 *   lse  fp32 [B, H, N]      (log-sum-exp of the scaled scores, natural log) saved for backward
 * Supported: N in {4, 16, 64, 256} (any N <= 256 with N % 4 == 0 on the generic path), hd % 8 == 0, hd <= 128.
 */
int tae_attention_fwd(const tae_bf16* qkv, tae_bf16* out, float* lse, int32_t B, int32_t N,
                      int32_t H, int32_t hd, void* stream);
/* dqkv bf16 [B*N, 3*H*hd] from dout bf16 [B*N, H*hd].  tae_attention_bwd_delta takes delta = rowsum(dout * out) per
 * (image, head, token) as fp32 [B, H, N] (e.g. from a TAE_EPI_BF16_ROWDOT GEMM) instead of `out`; supported for the
 * N = 256 and N = 64 kernels with hd = 64 (the output tile is then neither read nor staged). */
int tae_attention_bwd_delta(const tae_bf16* qkv, const tae_bf16* dout, const float* lse, const float* delta,
                            tae_bf16* dqkv, int32_t B, int32_t N, int32_t H, int32_t hd, void* stream);
int tae_attention_bwd(const tae_bf16* qkv, const tae_bf16* out, const tae_bf16* dout,
                      const float* lse, tae_bf16* dqkv, int32_t B, int32_t N, int32_t H, int32_t hd,
                      void* stream);

/* ---- Patch index maps (pure integer permutations, bit-exact) -------------------------------
 * im2col for PatchEmbed (tae.py:43,50-52): cols[b*N + (h*g+w), c*p*p + i*p + j] = bf16(imgs[b,c,h*p+i,w*p+j])
 * imgs fp32 NCHW [B, 3, S, S], S % p == 0, g = S/p, N = g*g; cols bf16 [B*N, 3*p*p].
 */
int tae_im2col_bf16(const float* imgs, tae_bf16* cols, int32_t B, int32_t S, int32_t p, void* stream);
/* TAE.patchify (tae.py:196-208): out[b, h*g+w, (i*p+j)*3 + c] = imgs[b,c,h*p+i,w*p+j]; elem_size in {2,4} bytes */
int tae_patchify(const void* imgs, void* out, int32_t B, int32_t S, int32_t p, int32_t elem_size, void* stream);
/* TAE.unpatchify (tae.py:210-222): exact inverse of tae_patchify */
int tae_unpatchify(const void* x, void* imgs, int32_t B, int32_t S, int32_t p, int32_t elem_size, void* stream);

/* C-channel variants for the downstream segmentation head: VITForSegmentation.unpatchify (tae.py:391-403),
 * out[b, c, h*p+i, w*p+j] = x[b, h*g+w, (i*p+j)*C + c], and its adjoint (the gradient's path back). */
int tae_patchify_c(const void* imgs, void* out, int32_t B, int32_t S, int32_t p, int32_t C, int32_t elem_size, void* stream);
int tae_unpatchify_c(const void* x, void* imgs, int32_t B, int32_t S, int32_t p, int32_t C, int32_t elem_size, void* stream);

/* ---- Loss ---------------------------------------------------------------------------------
 * TAE.forward_loss (tae.py:256-265) with patchify folded into the indexing:
 *   loss = mean over all B*N*3p^2 elements of (float(pred) - patchify(imgs))^2      (fp32)
 *   dpred = bf16( grad_scale * 2 * (pred - target) / numel )      (if dpred != NULL)
 * loss_accum: fp32[1], must be zeroed by the caller; receives sum/numel.
 * grad_scale: device pointer to fp32[1] (the upstream gradient, e.g. GradScaler scale) or NULL (= 1).
 */
int tae_mse_loss(const tae_bf16* pred, const float* imgs, float* loss_accum, tae_bf16* dpred,
                 const float* grad_scale, int32_t B, int32_t S, int32_t p, void* stream);

/* ---- Reductions for bias / pos-embed gradients ------------------------------------------------
 * out[n] (+)= sum_m float(x[m, n]),  x bf16 [M, N] row-major (ld = ldx).  Bias gradients.
 * workspace: fp32 [tae_colsum_workspace_floats(M, N)].
 */
size_t tae_colsum_workspace_floats(int32_t M, int32_t N);
int tae_colsum_bf16(const tae_bf16* x, int32_t M, int32_t N, int32_t ldx, float* out, int32_t accumulate,
                    float* workspace, void* stream);
/* out[n] (+)= sum_r x[r, n],  x fp32 [R, N] (e.g. the colsum_partials of tae_gemm). */
int tae_colsum_f32(const float* x, int32_t R, int32_t N, float* out, int32_t accumulate, void* stream);
/* out[r, :] (+)= sum_b x[b*R + r, :],  x fp32 [B*R, D].  pos_embed / decoder_pos_embed gradients. */
int tae_batch_sum_f32(const float* x, int32_t B, int32_t R, int32_t D, float* out, int32_t accumulate,
                      void* stream);

/* global average pooling over tokens, VITForRecognition.forward_head (tae.py:333): out[b,:] = mean_n x[b*N+n,:],
 * and its gradient dx[b*N+n,:] = dy[b,:] / N.  fp32, D % 4 == 0. */
int tae_token_mean_f32(const float* x, float* out, int32_t B, int32_t N, int32_t D, void* stream);
int tae_token_mean_bwd_f32(const float* dy, float* dx, int32_t B, int32_t N, int32_t D, void* stream);

/* ---- Optimizer ------------------------------------------------------------------------------
 * torch.optim.AdamW(fused=True) (train.py:109) over a flat fp32 arena, one launch per call:
 *   g' = g * grad_scale (1/world for DDP mean, 1/loss_scale for fp16)
 *   m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2
 *   p = p (1 - lr wd) - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * and, in the same pass, the bf16 shadow copy used by the GEMMs (p_bf16 may be NULL).
 * `step` is t (>= 1).  If grad_sq_sum != NULL, sum(g'^2) is atomically added to it (get_grad_norm_,
 * util/misc.py:274-286).  If found_inf != NULL and *found_inf != 0 on the device, the update is skipped
 * (GradScaler.step semantics, util/misc.py:261).
 */
int tae_adamw_step(float* p, const float* g, float* m, float* v, tae_bf16* p_bf16, size_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                   float grad_scale, float* grad_sq_sum, const int32_t* found_inf, void* stream);
/* The same step with its scalars in DEVICE memory, so that a CUDA graph of the whole training step (train.py:122-150)
 * can be replayed with a new learning rate (util/misc.py:400-412 writes it every iteration) and step count:
 * tae_adamw_hyper is a HOST function that fills `out[TAE_ADAMW_HYPER_FLOATS]` with the derived scalars (the very ones
 * tae_adamw_step computes internally — both paths are bit-identical); the caller copies them to the device and passes
 * that pointer as `hyper_dev`. */
#define TAE_ADAMW_HYPER_FLOATS 9
int tae_adamw_hyper(float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                    float grad_scale, float* out);
int tae_adamw_step_dev(float* p, const float* g, float* m, float* v, tae_bf16* p_bf16, size_t n,
                       const float* hyper_dev, float* grad_sq_sum, const int32_t* found_inf, void* stream);
/* dst = bf16(src), n elements */
int tae_cast_f32_to_bf16(const float* src, tae_bf16* dst, size_t n, void* stream);
/* non-finite check + sum of squares over a fp32 arena: found_inf[0] |= any(!isfinite), sq_sum[0] += sum g^2 */
int tae_grad_stats(const float* g, size_t n, float* sq_sum, int32_t* found_inf, void* stream);

/* ---- fp32 ("no autocast") mode ------------------------------------------------------------------
 * The reference without autocast keeps every activation in fp32 (tae.py run as plain nn.Module code).  In this
 * mode the GEMMs still run on the tcgen05 tensor cores: each fp32 operand is split into three bf16 terms
 * (x = hi + mid + lo, exact), and the six significant cross products are issued as tae_gemm launches that
 * accumulate into one fp32 output (TAE_EPI_F32_ACC, beta = 1) -> fp32-level error (dropped terms <= 2^-24).
 * The entry points below are the fp32 element-wise / LayerNorm / attention / loss pieces around those GEMMs
 * (same reference call sites as their bf16 counterparts above).
 */
int tae_split3_bf16(const float* x, tae_bf16* hi, tae_bf16* mid, tae_bf16* lo, size_t n, void* stream);
/* y[m,n] += bias[n] + resid[m % resid_rows, n] (either may be NULL); act (optional) = gelu_erf(y)  (tae.py:101-104,129-130) */
int tae_bias_act_f32(float* y, const float* bias, const float* resid, int32_t resid_rows, float* act,
                     int32_t M, int32_t N, void* stream);
/* dh = da * gelu_erf'(h) */
int tae_gelu_bwd_f32(const float* h, const float* da, float* dh, size_t n, void* stream);
/* out = a + b (b may be NULL) */
int tae_add_f32(const float* a, const float* b, float* out, size_t n, void* stream);
int tae_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                          float* rstd, int32_t rows, int32_t D, float eps, void* stream);
/* dres_out = dres_in + LN'(dy); dgamma/dbeta (optional) written or accumulated per bit 0/1 of `accumulate` */
int tae_layernorm_bwd_f32(const float* dy, const float* x, const float* mean, const float* rstd,
                          const float* gamma, const float* dres_in, float* dres_out, float* dgamma,
                          float* dbeta, int32_t accumulate, int32_t rows, int32_t D, void* stream);
/* same tensor layouts as tae_attention_fwd/bwd with fp32 elements; workspace fp32 [2*B*H*N*N] */
int tae_attention_fwd_f32(const float* qkv, float* out, float* lse, int32_t B, int32_t N, int32_t H,
                          int32_t hd, void* stream);
size_t tae_attention_bwd_f32_workspace_floats(int32_t B, int32_t N, int32_t H);
int tae_attention_bwd_f32(const float* qkv, const float* out, const float* dout, const float* lse,
                          float* dqkv, float* workspace, int32_t B, int32_t N, int32_t H, int32_t hd,
                          void* stream);
int tae_im2col_f32(const float* imgs, float* cols, int32_t B, int32_t S, int32_t p, void* stream);
int tae_mse_loss_f32(const float* pred, const float* imgs, float* loss_accum, float* dpred,
                     const float* grad_scale, int32_t B, int32_t S, int32_t p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAE_B200_H_ */
