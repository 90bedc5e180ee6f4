"""Tensor-level wrappers over the C ABI (include/tae_b200.h).  PyTorch only supplies device memory and streams.

Every function launches asynchronously on the current CUDA stream and returns torch tensors that own the
output memory.  Nothing here has a CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_BF16_ROWDOT, EPI_F32_ACC, EPI_F32_RESID, GemmArgs, check)

bf16 = torch.bfloat16
f32 = torch.float32


def _L():
    return _lib.load()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.TaeError(f"{name}: expected a CUDA tensor (tae_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.TaeError(f"{name}: expected dtype {dtype}, got {t.dtype}")


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def gemm(A: torch.Tensor, B: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, epilogue: int = EPI_BF16,
         out: torch.Tensor | None = None, out2: torch.Tensor | None = None, bias: torch.Tensor | None = None,
         resid: torch.Tensor | None = None, resid_rows: int = 0, aux: torch.Tensor | None = None, beta: int = 0,
         splits: int = 0, colsum_partials: torch.Tensor | None = None, rowdot_tokens: int = 0,
         gelu_grad: bool = True):
    """D[M,N] = sum_k A(m,k) B(n,k) on the tcgen05 tensor cores; see include/tae_b200.h for the conventions.

    A: [M,K] (or [K,M] if a_mn); B: [N,K] (or [K,N] if b_mn); 2-D bf16, inner stride 1.
    """
    _req(A, bf16, "gemm A")
    _req(B, bf16, "gemm B")
    assert A.dim() == 2 and B.dim() == 2 and A.stride(1) == 1 and B.stride(1) == 1
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    if b_mn:
        Kb, N = B.shape
    else:
        N, Kb = B.shape
    if K != Kb:
        raise _lib.TaeError(f"gemm: reduction dims differ (A gives K={K}, B gives K={Kb})")
    out_dtype = f32 if epilogue in (EPI_F32_RESID, EPI_F32_ACC) else bf16
    if epilogue == EPI_BF16_GELU and not gelu_grad:
        # inference: only gelu(h) is produced; returns (None, gelu(h))
        if out2 is None:
            out2 = torch.empty((M, N), dtype=bf16, device=A.device)
        _req(out2, bf16, "gemm out2")
        assert out2.dim() == 2 and out2.shape == (M, N) and out2.stride(1) == 1
        a = GemmArgs()
        a.A, a.B, a.M, a.N, a.K = A.data_ptr(), B.data_ptr(), M, N, K
        a.lda, a.ldb, a.a_mn_major, a.b_mn_major = A.stride(0), B.stride(0), int(a_mn), int(b_mn)
        a.epilogue, a.out, a.ldo, a.out2 = epilogue, 0, out2.stride(0), out2.data_ptr()
        if bias is not None:
            _req(bias, f32, "gemm bias")
            assert bias.numel() == N
            a.bias = bias.data_ptr()
        check(_L().tae_gemm(C.byref(a), _stream()), "tae_gemm")
        return None, out2
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=A.device)
    else:
        _req(out, out_dtype, "gemm out")
        assert out.dim() == 2 and out.shape == (M, N) and out.stride(1) == 1
    a = GemmArgs()
    a.A, a.B = A.data_ptr(), B.data_ptr()
    a.M, a.N, a.K = M, N, K
    a.lda, a.ldb = A.stride(0), B.stride(0)
    a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
    a.epilogue = epilogue
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    if epilogue == EPI_BF16_GELU:
        if out2 is None:
            out2 = torch.empty((M, N), dtype=bf16, device=A.device)
        assert out2.stride(0) == out.stride(0) and out2.dtype == bf16
        a.out2 = out2.data_ptr()
    if bias is not None:
        _req(bias, f32, "gemm bias")
        assert bias.numel() == N
        a.bias = bias.data_ptr()
    if epilogue == EPI_F32_RESID:
        _req(resid, f32, "gemm resid")
        assert resid.dim() == 2 and resid.shape[1] == N and resid.stride(1) == 1
        a.resid, a.ldr = resid.data_ptr(), resid.stride(0)
        a.resid_rows = resid_rows if resid_rows > 0 else resid.shape[0]
    rowdot = None
    if epilogue == EPI_BF16_ROWDOT:
        assert rowdot_tokens > 0 and M % rowdot_tokens == 0 and N % 64 == 0
        rowdot = torch.empty((M // rowdot_tokens, N // 64, rowdot_tokens), dtype=f32, device=A.device)
        a.rowdot, a.rowdot_tokens = rowdot.data_ptr(), rowdot_tokens
    if epilogue in (EPI_BF16_DGELU, EPI_BF16_ROWDOT):
        _req(aux, bf16, "gemm aux")
        assert aux.shape == (M, N) and aux.stride(1) == 1
        a.aux, a.ldaux = aux.data_ptr(), aux.stride(0)
    a.beta = int(beta)
    a.splits = int(splits)
    if colsum_partials is not None:
        _req(colsum_partials, f32, "gemm colsum_partials")
        assert colsum_partials.is_contiguous() and colsum_partials.shape == ((M + 31) // 32, N)
        a.colsum_partials = colsum_partials.data_ptr()
    check(_L().tae_gemm(C.byref(a), _stream()), "tae_gemm")
    if epilogue == EPI_BF16_GELU:
        return out, out2
    if epilogue == EPI_BF16_ROWDOT:
        return out, rowdot
    return out


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float):
    """x fp32 [rows, D] -> (y bf16 [rows, D], mean fp32 [rows], rstd fp32 [rows])."""
    _req(x, f32, "layernorm x")
    assert x.dim() == 2 and x.is_contiguous()
    rows, D = x.shape
    y = torch.empty((rows, D), dtype=bf16, device=x.device)
    stats = torch.empty((2, rows), dtype=f32, device=x.device)
    check(_L().tae_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), stats[0].data_ptr(),
                                 stats[1].data_ptr(), rows, D, float(eps), _stream()), "tae_layernorm_fwd")
    return y, stats[0], stats[1]


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
                  dres_in: torch.Tensor | None, *, want_bf16: bool = True, dgamma=None, dbeta=None, dcolsum=None,
                  acc_mask: int = 0):
    """LayerNorm backward + residual-gradient add + bf16 re-cast.

    Returns (dres_out fp32, dres_out_bf16 | None, dgamma, dbeta, dcolsum | None).  dgamma/dbeta/dcolsum may be given
    (e.g. views into a gradient arena); acc_mask bit 0/1/2 selects `+=` for dgamma/dbeta/dcolsum.
    """
    _req(dy, bf16, "layernorm_bwd dy")
    _req(x, f32, "layernorm_bwd x")
    rows, D = x.shape
    L = _L()
    nparts = L.tae_layernorm_bwd_num_partials(rows, D)
    if nparts <= 0:
        raise _lib.TaeError(f"layernorm_bwd: unsupported shape rows={rows} D={D}")
    partials = torch.empty((nparts, 3, D), dtype=f32, device=x.device)
    dres_out = torch.empty((rows, D), dtype=f32, device=x.device)
    dres_b = torch.empty((rows, D), dtype=bf16, device=x.device) if want_bf16 else None
    check(L.tae_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                              _ptr(dres_in), dres_out.data_ptr(), _ptr(dres_b), partials.data_ptr(), rows, D,
                              _stream()), "tae_layernorm_bwd")
    if dgamma is None:
        dgamma = torch.empty((D,), dtype=f32, device=x.device)
        acc_mask &= ~1
    if dbeta is None:
        dbeta = torch.empty((D,), dtype=f32, device=x.device)
        acc_mask &= ~2
    if dcolsum is None and want_bf16:
        dcolsum = torch.empty((D,), dtype=f32, device=x.device)
        acc_mask &= ~4
    check(L.tae_layernorm_bwd_finalize(partials.data_ptr(), nparts, D, dgamma.data_ptr(), dbeta.data_ptr(),
                                       _ptr(dcolsum), acc_mask, _stream()), "tae_layernorm_bwd_finalize")
    return dres_out, dres_b, dgamma, dbeta, dcolsum


# ------------------------------------------------------------------------------------------------
# Attention
# ------------------------------------------------------------------------------------------------
def attention_fwd(qkv: torch.Tensor, B: int, N: int, H: int, hd: int):
    """qkv bf16 [B*N, 3*H*hd] -> (out bf16 [B*N, H*hd], lse fp32 [B, H, N])."""
    _req(qkv, bf16, "attention qkv")
    assert qkv.is_contiguous() and qkv.shape == (B * N, 3 * H * hd)
    out = torch.empty((B * N, H * hd), dtype=bf16, device=qkv.device)
    lse = torch.empty((B, H, N), dtype=f32, device=qkv.device)
    check(_L().tae_attention_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, hd, _stream()),
          "tae_attention_fwd")
    return out, lse


def attention_takes_delta(N: int, hd: int) -> bool:
    """True when attention_bwd accepts delta = rowsum(dout * out) precomputed by a TAE_EPI_BF16_ROWDOT GEMM."""
    return N in (256, 64) and hd == 64


def attention_bwd(qkv, out, dout, lse, B: int, N: int, H: int, hd: int, delta: torch.Tensor | None = None):
    _req(dout, bf16, "attention dout")
    assert dout.is_contiguous() and qkv.is_contiguous() and (out is None or out.is_contiguous())
    dqkv = torch.empty_like(qkv)
    if delta is not None:
        _req(delta, f32, "attention delta")
        assert delta.is_contiguous() and delta.shape == (B, H, N)
        check(_L().tae_attention_bwd_delta(qkv.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                           B, N, H, hd, _stream()), "tae_attention_bwd_delta")
        return dqkv
    check(_L().tae_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, N,
                                 H, hd, _stream()), "tae_attention_bwd")
    return dqkv


# ------------------------------------------------------------------------------------------------
# Patch index maps, loss
# ------------------------------------------------------------------------------------------------
def im2col(imgs: torch.Tensor, p: int) -> torch.Tensor:
    """imgs fp32 [B,3,S,S] -> cols bf16 [B*(S/p)^2, 3*p*p] in (c,i,j) order (the conv weight's layout)."""
    _req(imgs, f32, "im2col imgs")
    assert imgs.dim() == 4 and imgs.is_contiguous() and imgs.shape[1] == 3 and imgs.shape[2] == imgs.shape[3]
    B, _, S, _ = imgs.shape
    g = S // p
    cols = torch.empty((B * g * g, 3 * p * p), dtype=bf16, device=imgs.device)
    check(_L().tae_im2col_bf16(imgs.data_ptr(), cols.data_ptr(), B, S, p, _stream()), "tae_im2col_bf16")
    return cols


def patchify(imgs: torch.Tensor, p: int) -> torch.Tensor:
    """TAE.patchify (tae.py:196-208): [B,3,S,S] -> [B, (S/p)^2, 3*p*p], dtype preserved (2- or 4-byte types)."""
    if not imgs.is_cuda:
        raise _lib.TaeError("patchify: expected a CUDA tensor (tae_b200 has no CPU fallback)")
    assert imgs.dim() == 4 and imgs.shape[1] == 3 and imgs.shape[2] == imgs.shape[3] and imgs.shape[2] % p == 0
    imgs = imgs.contiguous()
    es = imgs.element_size()
    assert es in (2, 4), "patchify supports 2- and 4-byte element types"
    B, _, S, _ = imgs.shape
    g = S // p
    out = torch.empty((B, g * g, 3 * p * p), dtype=imgs.dtype, device=imgs.device)
    check(_L().tae_patchify(imgs.data_ptr(), out.data_ptr(), B, S, p, es, _stream()), "tae_patchify")
    return out


def unpatchify(x: torch.Tensor, p: int) -> torch.Tensor:
    """TAE.unpatchify (tae.py:210-222): [B, L, 3*p*p] -> [B,3,S,S] with S = sqrt(L)*p."""
    if not x.is_cuda:
        raise _lib.TaeError("unpatchify: expected a CUDA tensor (tae_b200 has no CPU fallback)")
    g = int(x.shape[1] ** .5)
    assert g * g == x.shape[1] and x.shape[2] == 3 * p * p
    x = x.contiguous()
    es = x.element_size()
    assert es in (2, 4)
    B = x.shape[0]
    S = g * p
    imgs = torch.empty((B, 3, S, S), dtype=x.dtype, device=x.device)
    check(_L().tae_unpatchify(x.data_ptr(), imgs.data_ptr(), B, S, p, es, _stream()), "tae_unpatchify")
    return imgs


def unpatchify_c(x: torch.Tensor, p: int, C: int) -> torch.Tensor:
    """VITForSegmentation.unpatchify (tae.py:391-403): [B, L, p*p*C] -> [B, C, S, S]."""
    if not x.is_cuda:
        raise _lib.TaeError("unpatchify_c: expected a CUDA tensor (tae_b200 has no CPU fallback)")
    g = int(x.shape[1] ** .5)
    assert g * g == x.shape[1] and x.shape[2] == C * p * p
    x = x.contiguous()
    es = x.element_size()
    assert es in (2, 4)
    B, S = x.shape[0], g * p
    imgs = torch.empty((B, C, S, S), dtype=x.dtype, device=x.device)
    check(_L().tae_unpatchify_c(x.data_ptr(), imgs.data_ptr(), B, S, p, C, es, _stream()), "tae_unpatchify_c")
    return imgs


def patchify_c(imgs: torch.Tensor, p: int) -> torch.Tensor:
    """Adjoint / inverse of unpatchify_c: [B, C, S, S] -> [B, (S/p)^2, p*p*C]."""
    if not imgs.is_cuda:
        raise _lib.TaeError("patchify_c: expected a CUDA tensor (tae_b200 has no CPU fallback)")
    imgs = imgs.contiguous()
    es = imgs.element_size()
    assert es in (2, 4)
    B, Cc, S, _ = imgs.shape
    g = S // p
    out = torch.empty((B, g * g, Cc * p * p), dtype=imgs.dtype, device=imgs.device)
    check(_L().tae_patchify_c(imgs.data_ptr(), out.data_ptr(), B, S, p, Cc, es, _stream()), "tae_patchify_c")
    return out


def token_mean(x: torch.Tensor, B: int, N: int) -> torch.Tensor:
    """fp32 [B*N, D] -> fp32 [B, D]: mean over the tokens of each image (tae.py:333)."""
    _req(x, f32, "token_mean x")
    assert x.is_contiguous()
    D = x.shape[-1]
    out = torch.empty((B, D), dtype=f32, device=x.device)
    check(_L().tae_token_mean_f32(x.data_ptr(), out.data_ptr(), B, N, D, _stream()), "tae_token_mean_f32")
    return out


def token_mean_bwd(dy: torch.Tensor, B: int, N: int) -> torch.Tensor:
    _req(dy, f32, "token_mean_bwd dy")
    dy = dy.contiguous()
    D = dy.shape[-1]
    dx = torch.empty((B * N, D), dtype=f32, device=dy.device)
    check(_L().tae_token_mean_bwd_f32(dy.data_ptr(), dx.data_ptr(), B, N, D, _stream()), "tae_token_mean_bwd_f32")
    return dx


def mse_loss(pred: torch.Tensor, imgs: torch.Tensor, p: int, *, want_grad: bool = False,
             grad_scale: torch.Tensor | None = None):
    """forward_loss (tae.py:256-265).  Returns (loss fp32 0-dim, dpred bf16 | None)."""
    _req(pred, bf16, "mse_loss pred")
    _req(imgs, f32, "mse_loss imgs")
    assert pred.is_contiguous() and imgs.is_contiguous()
    B, _, S, _ = imgs.shape
    loss = torch.zeros((), dtype=f32, device=pred.device)
    dpred = torch.empty_like(pred) if want_grad else None
    check(_L().tae_mse_loss(pred.data_ptr(), imgs.data_ptr(), loss.data_ptr(), _ptr(dpred), _ptr(grad_scale), B, S, p,
                            _stream()), "tae_mse_loss")
    return loss, dpred


# ------------------------------------------------------------------------------------------------
# Reductions
# ------------------------------------------------------------------------------------------------
def colsum(x: torch.Tensor, out: torch.Tensor | None = None, accumulate: bool = False) -> torch.Tensor:
    """sum over rows of a bf16 [M,N] matrix -> fp32 [N]  (bias gradients)."""
    _req(x, bf16, "colsum x")
    assert x.dim() == 2 and x.stride(1) == 1
    M, N = x.shape
    L = _L()
    ws = torch.empty((L.tae_colsum_workspace_floats(M, N),), dtype=f32, device=x.device)
    if out is None:
        out = torch.empty((N,), dtype=f32, device=x.device)
        accumulate = False
    check(L.tae_colsum_bf16(x.data_ptr(), M, N, x.stride(0), out.data_ptr(), int(accumulate), ws.data_ptr(), _stream()),
          "tae_colsum_bf16")
    return out


def colsum_f32(x: torch.Tensor, out: torch.Tensor | None = None, accumulate: bool = False) -> torch.Tensor:
    """sum over rows of a small fp32 [R,N] matrix (the per-32-row partial sums a GEMM epilogue emitted) -> fp32 [N]."""
    _req(x, f32, "colsum_f32 x")
    assert x.dim() == 2 and x.is_contiguous()
    R, N = x.shape
    if out is None:
        out = torch.empty((N,), dtype=f32, device=x.device)
        accumulate = False
    check(_L().tae_colsum_f32(x.data_ptr(), R, N, out.data_ptr(), int(accumulate), _stream()), "tae_colsum_f32")
    return out


def batch_sum(x: torch.Tensor, B: int, R: int, out: torch.Tensor | None = None, accumulate: bool = False):
    """x fp32 [B*R, D] -> fp32 [R, D] summed over the batch (pos-embed gradients)."""
    _req(x, f32, "batch_sum x")
    assert x.is_contiguous()
    D = x.shape[-1]
    if out is None:
        out = torch.empty((R, D), dtype=f32, device=x.device)
        accumulate = False
    check(_L().tae_batch_sum_f32(x.data_ptr(), B, R, D, out.data_ptr(), int(accumulate), _stream()), "tae_batch_sum_f32")
    return out


# ------------------------------------------------------------------------------------------------
# Optimizer / casts
# ------------------------------------------------------------------------------------------------
def cast_bf16(src: torch.Tensor, dst: torch.Tensor | None = None) -> torch.Tensor:
    _req(src, f32, "cast src")
    assert src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=bf16, device=src.device)
    check(_L().tae_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "tae_cast_f32_to_bf16")
    return dst


def adamw_step(p, g, m, v, p_bf16, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, grad_sq_sum=None,
               found_inf=None):
    n = p.numel()
    check(_L().tae_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(p_bf16), n, float(lr),
                              float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale),
                              _ptr(grad_sq_sum), _ptr(found_inf), _stream()), "tae_adamw_step")


ADAMW_HYPER_FLOATS = 9


def adamw_hyper(*, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0) -> list:
    """Host side of the device-scalar AdamW step: the 9 derived floats `tae_adamw_step_dev` reads from device memory."""
    buf = (C.c_float * ADAMW_HYPER_FLOATS)()
    check(_L().tae_adamw_hyper(float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                               float(grad_scale), buf), "tae_adamw_hyper")
    return list(buf)


def adamw_step_dev(p, g, m, v, p_bf16, hyper_dev: torch.Tensor, *, grad_sq_sum=None, found_inf=None):
    """AdamW step whose scalars come from `hyper_dev` (fp32 [9] on the device): replayable inside a CUDA graph."""
    _req(hyper_dev, f32, "adamw hyper")
    assert hyper_dev.numel() == ADAMW_HYPER_FLOATS and hyper_dev.is_contiguous()
    check(_L().tae_adamw_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(p_bf16), p.numel(),
                                  hyper_dev.data_ptr(), _ptr(grad_sq_sum), _ptr(found_inf), _stream()), "tae_adamw_step_dev")


def grad_stats(g: torch.Tensor, sq_sum: torch.Tensor | None, found_inf: torch.Tensor | None):
    check(_L().tae_grad_stats(g.data_ptr(), g.numel(), _ptr(sq_sum), _ptr(found_inf), _stream()), "tae_grad_stats")


def set_dynamic_scheduling(enable: bool) -> int:
    """Dynamic work lists for the persistent kernels (see include/tae_b200.h); returns the previous setting."""
    return int(_L().tae_set_dynamic_scheduling(int(bool(enable))))


def launch_count() -> int:
    return int(_L().tae_launch_count())
