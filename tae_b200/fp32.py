"""fp32 ("no autocast") numerics mode of the TAE hot path: the reference run as plain fp32 nn.Module code
(tae.py:46-54, 72-82, 100-105, 128-131, 224-271), gated at 1e-4 relative error (BASELINE.json north star).

Every activation stays in fp32.  GEMMs still run on the tcgen05 tensor cores: each fp32 operand is split into three
bf16 terms (x = hi + mid + lo exactly) and the six significant cross products are accumulated into one fp32 output
by six `tae_gemm` launches (TAE_EPI_F32_ACC, beta=1), smallest terms first.  bf16 x bf16 products are exact in
fp32, so the result carries fp32-level rounding error only (dropped terms <= 2^-24 relative).  LayerNorm, GELU,
attention, im2col and the loss use the fp32 kernels of csrc/fp32_mode.cu.

The autograd structure mirrors tae_b200/tae.py: one Function per transformer block etc., all backward passes are
hand-written sequences of C-ABI calls (no ATen arithmetic on the path).  Select with `model.set_precision("fp32")`.
This mode exists for parity checking; it is ~10x slower than the bf16 path and is not the benchmarked one.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import EPI_F32_ACC, check

f32 = torch.float32
bf16 = torch.bfloat16
_L = ops._L
_stream = ops._stream
_ptr = ops._ptr


# ----------------------------------------------------------------------------------------------------
# op wrappers
# ----------------------------------------------------------------------------------------------------
def split3(x: torch.Tensor):
    """fp32 [R, C] -> three bf16 [R, C] tensors with hi + mid + lo == x (exactly, barring underflow)."""
    ops._req(x, f32, "split3 x")
    x = x.contiguous()
    parts = torch.empty((3,) + tuple(x.shape), dtype=bf16, device=x.device)
    check(_L().tae_split3_bf16(x.data_ptr(), parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr(), x.numel(),
                               _stream()), "tae_split3_bf16")
    return parts[0], parts[1], parts[2]


# (index into A's parts, index into B's parts), smallest magnitude first
_TERMS = ((2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0))


def gemm_f32(A3, B3, *, a_mn: bool = False, b_mn: bool = False, out: torch.Tensor | None = None, beta: int = 0):
    """fp32-accurate D = A B^T from 3-way bf16 splits: six tensor-core GEMMs accumulating into one fp32 output."""
    first = True
    for ia, ib in _TERMS:
        out = ops.gemm(A3[ia], B3[ib], a_mn=a_mn, b_mn=b_mn, epilogue=EPI_F32_ACC, out=out,
                       beta=(beta if first else 1), splits=1)
        first = False
    return out


def bias_act(y: torch.Tensor, bias=None, resid=None, resid_rows: int = 0, want_act: bool = False):
    """In place y += bias + resid[row % resid_rows]; returns gelu_erf(y) as a new tensor if want_act."""
    M, N = y.shape
    act = torch.empty_like(y) if want_act else None
    rr = 0 if resid is None else (resid_rows if resid_rows > 0 else resid.shape[0])
    check(_L().tae_bias_act_f32(y.data_ptr(), _ptr(bias), _ptr(resid), rr, _ptr(act), M, N, _stream()), "tae_bias_act_f32")
    return act


def gelu_bwd(h, da):
    dh = torch.empty_like(h)
    check(_L().tae_gelu_bwd_f32(h.data_ptr(), da.data_ptr(), dh.data_ptr(), h.numel(), _stream()), "tae_gelu_bwd_f32")
    return dh


def layernorm_fwd(x, gamma, beta, eps):
    rows, D = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((2, rows), dtype=f32, device=x.device)
    check(_L().tae_layernorm_fwd_f32(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), stats[0].data_ptr(),
                                     stats[1].data_ptr(), rows, D, float(eps), _stream()), "tae_layernorm_fwd_f32")
    return y, stats[0], stats[1]


def layernorm_bwd(dy, x, mean, rstd, gamma, dres_in, dgamma=None, dbeta=None, acc_mask: int = 0):
    rows, D = x.shape
    dres_out = torch.empty_like(x)
    if dgamma is None:
        dgamma = torch.empty((D,), dtype=f32, device=x.device)
        acc_mask &= ~1
    if dbeta is None:
        dbeta = torch.empty((D,), dtype=f32, device=x.device)
        acc_mask &= ~2
    check(_L().tae_layernorm_bwd_f32(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                     _ptr(dres_in), dres_out.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), acc_mask,
                                     rows, D, _stream()), "tae_layernorm_bwd_f32")
    return dres_out, dgamma, dbeta


def attention_fwd(qkv, B, N, H, hd):
    out = torch.empty((B * N, H * hd), dtype=f32, device=qkv.device)
    lse = torch.empty((B, H, N), dtype=f32, device=qkv.device)
    check(_L().tae_attention_fwd_f32(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, hd, _stream()),
          "tae_attention_fwd_f32")
    return out, lse


def attention_bwd(qkv, out, dout, lse, B, N, H, hd):
    dqkv = torch.empty_like(qkv)
    ws = torch.empty((_L().tae_attention_bwd_f32_workspace_floats(B, N, H),), dtype=f32, device=qkv.device)
    check(_L().tae_attention_bwd_f32(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                     ws.data_ptr(), B, N, H, hd, _stream()), "tae_attention_bwd_f32")
    return dqkv


def im2col(imgs, p):
    B, _, S, _ = imgs.shape
    g = S // p
    cols = torch.empty((B * g * g, 3 * p * p), dtype=f32, device=imgs.device)
    check(_L().tae_im2col_f32(imgs.data_ptr(), cols.data_ptr(), B, S, p, _stream()), "tae_im2col_f32")
    return cols


def mse_loss(pred, imgs, p, want_grad=False, grad_scale=None):
    B, _, S, _ = imgs.shape
    loss = torch.zeros((), dtype=f32, device=pred.device)
    dpred = torch.empty_like(pred) if want_grad else None
    check(_L().tae_mse_loss_f32(pred.data_ptr(), imgs.data_ptr(), loss.data_ptr(), _ptr(dpred), _ptr(grad_scale), B, S, p,
                                _stream()), "tae_mse_loss_f32")
    return loss, dpred


# ----------------------------------------------------------------------------------------------------
# building blocks shared by the autograd functions
# ----------------------------------------------------------------------------------------------------
def _linear_fwd(x, w, b=None, resid=None, resid_rows=0):
    """y = x W^T (+ b) (+ resid): fp32 [M,K] x [N,K] -> fp32 [M,N]."""
    y = gemm_f32(split3(x), split3(w.detach().reshape(w.shape[0], -1)))
    if b is not None or resid is not None:
        bias_act(y, None if b is None else b.detach(), resid, resid_rows)
    return y


def _linear_bwd(dy, x, lin_w, lin_b, need_w, need_b, need_x):
    """Returns (dx | None, what autograd gets for W, for b).  Parameter gradients go straight to the fused optimizer's
    arena when it exists (tae._sink), exactly as in the bf16 path."""
    from .tae import _done, _sink

    dy3 = split3(dy)
    g_b = None
    if lin_b is not None and need_b:
        t, acc = _sink(lin_b)
        g_b = _done(lin_b, ops.colsum_f32(dy, out=t, accumulate=bool(acc)))
    g_w = None
    if need_w:
        t, acc = _sink(lin_w)
        out = None if t is None else t.view(dy.shape[1], x.shape[1])
        g = gemm_f32(dy3, split3(x), a_mn=True, b_mn=True, out=out, beta=acc)
        g_w = _done(lin_w, g.view(lin_w.shape))
    dx = None
    if need_x:
        dx = gemm_f32(dy3, split3(lin_w.detach().reshape(lin_w.shape[0], -1)), b_mn=True)
    return dx, g_w, g_b


def _f32_2d(t, D):
    t = t.reshape(-1, D)
    if t.dtype != f32:
        t = t.float()
    return t.contiguous()


def _pos_grad(pos, dx, B, N, D):
    from .tae import _done, _sink

    t, acc = _sink(pos)
    g = ops.batch_sum(dx, B, N, out=None if t is None else t.view(N, D), accumulate=bool(acc))
    return _done(pos, g.view(1, N, D))


# ----------------------------------------------------------------------------------------------------
# autograd functions (same decomposition as tae_b200/tae.py)
# ----------------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    """PatchEmbed conv as im2col GEMM + bias + pos_embed  (tae.py:46-54, :229)."""

    @staticmethod
    def forward(ctx, imgs, w, b, pos, mod):
        p = mod.patch_size[0]
        B, N, D = imgs.shape[0], mod.num_patches, w.shape[0]
        cols = im2col(imgs, p)
        x = _linear_fwd(cols, w, b, resid=pos.detach().view(N, D), resid_rows=N)
        ctx.mod, ctx.dims = mod, (B, N, D)
        ctx.save_for_backward(cols)
        return x.view(B, N, D)

    @staticmethod
    def backward(ctx, dx):
        (cols,) = ctx.saved_tensors
        B, N, D = ctx.dims
        proj = ctx.mod.proj
        need = ctx.needs_input_grad
        dx = _f32_2d(dx, D)
        _, gw, gb = _linear_bwd(dx, cols, proj.weight, proj.bias, need[1], need[2], False)
        gpos = _pos_grad(ctx.mod._pos_param, dx, B, N, D) if need[3] else None
        return None, gw, gb, gpos, None


class BlockFn(torch.autograd.Function):
    """x + attn(norm1(x)); x + mlp(norm2(x))  (tae.py:128-131), fp32 throughout."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkvw, qkvb, pw, pb, n2w, n2b, f1w, f1b, f2w, f2b, blk):
        B, N, D = x.shape
        H = blk.attn.num_heads
        hd = D // H
        x2 = _f32_2d(x, D)
        ln1, mean1, rstd1 = layernorm_fwd(x2, n1w.detach(), n1b.detach(), blk.norm1.eps)
        qkv = _linear_fwd(ln1, qkvw, qkvb)
        att, lse = attention_fwd(qkv, B, N, H, hd)
        xm = _linear_fwd(att, pw, pb, resid=x2)
        ln2, mean2, rstd2 = layernorm_fwd(xm, n2w.detach(), n2b.detach(), blk.norm2.eps)
        h = _linear_fwd(ln2, f1w, f1b)
        a = bias_act(h, want_act=True)
        xo = _linear_fwd(a, f2w, f2b, resid=xm)
        ctx.blk, ctx.dims = blk, (B, N, D, H, hd)
        ctx.save_for_backward(x2, mean1, rstd1, ln1, qkv, att, lse, xm, mean2, rstd2, ln2, h, a)
        return xo.view(B, N, D)

    @staticmethod
    def backward(ctx, dxo):
        from .tae import _done, _sink

        x2, mean1, rstd1, ln1, qkv, att, lse, xm, mean2, rstd2, ln2, h, a = ctx.saved_tensors
        blk = ctx.blk
        B, N, D, H, hd = ctx.dims
        need = ctx.needs_input_grad
        attn, mlp = blk.attn, blk.mlp
        dres = _f32_2d(dxo, D)
        # MLP branch
        da, g_f2w, g_f2b = _linear_bwd(dres, a, mlp.fc2.weight, mlp.fc2.bias, need[11], need[12], True)
        dh = gelu_bwd(h, da)
        dln2, g_f1w, g_f1b = _linear_bwd(dh, ln2, mlp.fc1.weight, mlp.fc1.bias, need[9], need[10], True)
        t_g, a_g = _sink(blk.norm2.weight)
        t_b, a_b = _sink(blk.norm2.bias)
        dres2, dg2, db2 = layernorm_bwd(dln2, xm, mean2, rstd2, blk.norm2.weight.detach(), dres, t_g, t_b, a_g | (a_b << 1))
        g_n2w, g_n2b = _done(blk.norm2.weight, dg2), _done(blk.norm2.bias, db2)
        # attention branch
        datt, g_pw, g_pb = _linear_bwd(dres2, att, attn.proj.weight, attn.proj.bias, need[5], need[6], True)
        dqkv = attention_bwd(qkv, att, datt, lse, B, N, H, hd)
        dln1, g_qw, g_qb = _linear_bwd(dqkv, ln1, attn.qkv.weight, attn.qkv.bias, need[3], need[4], True)
        t_g, a_g = _sink(blk.norm1.weight)
        t_b, a_b = _sink(blk.norm1.bias)
        dx, dg1, db1 = layernorm_bwd(dln1, x2, mean1, rstd1, blk.norm1.weight.detach(), dres2, t_g, t_b, a_g | (a_b << 1))
        g_n1w, g_n1b = _done(blk.norm1.weight, dg1), _done(blk.norm1.bias, db1)
        return (dx.view(B, N, D), g_n1w, g_n1b, g_qw, g_qb, g_pw, g_pb, g_n2w, g_n2b, g_f1w, g_f1b, g_f2w, g_f2b, None)


class NormLinearFn(torch.autograd.Function):
    """norm -> dict_proj (tae.py:234-237) and decoder_norm -> decoder_pred (tae.py:250-253)."""

    @staticmethod
    def forward(ctx, x, nw, nb, w, b, norm, lin):
        B, N, D = x.shape
        x2 = _f32_2d(x, D)
        ln, mean, rstd = layernorm_fwd(x2, nw.detach(), nb.detach(), norm.eps)
        y = _linear_fwd(ln, w, b)
        ctx.norm, ctx.lin, ctx.dims = norm, lin, (B, N, D)
        ctx.save_for_backward(x2, mean, rstd, ln)
        return y.view(B, N, -1)

    @staticmethod
    def backward(ctx, dy):
        from .tae import _done, _sink

        x2, mean, rstd, ln = ctx.saved_tensors
        norm, lin = ctx.norm, ctx.lin
        B, N, D = ctx.dims
        need = ctx.needs_input_grad
        dy2 = _f32_2d(dy, dy.shape[-1])
        dln, g_w, g_b = _linear_bwd(dy2, ln, lin.weight, lin.bias, need[3], need[4], True)
        t_g, a_g = _sink(norm.weight)
        t_b, a_b = _sink(norm.bias)
        dx, dg, db = layernorm_bwd(dln, x2, mean, rstd, norm.weight.detach(), None, t_g, t_b, a_g | (a_b << 1))
        return dx.view(B, N, D), _done(norm.weight, dg), _done(norm.bias, db), g_w, g_b, None, None


class EmbedLatentFn(torch.autograd.Function):
    """decoder_embed + decoder_pos_embed  (tae.py:242-245)."""

    @staticmethod
    def forward(ctx, z, w, b, pos, mod):
        B, N, V = z.shape
        D = w.shape[0]
        z2 = _f32_2d(z, V)
        x = _linear_fwd(z2, w, b, resid=pos.detach().view(N, D), resid_rows=N)
        ctx.mod, ctx.dims, ctx.zdtype = mod, (B, N, V, D), z.dtype
        ctx.save_for_backward(z2)
        return x.view(B, N, D)

    @staticmethod
    def backward(ctx, dx):
        (z2,) = ctx.saved_tensors
        mod = ctx.mod
        B, N, V, D = ctx.dims
        need = ctx.needs_input_grad
        lin = mod.decoder_embed
        dx = _f32_2d(dx, D)
        dz, g_w, g_b = _linear_bwd(dx, z2, lin.weight, lin.bias, need[1], need[2], need[0])
        g_pos = _pos_grad(mod.decoder_pos_embed, dx, B, N, D) if need[3] else None
        if dz is not None:
            dz = dz.view(B, N, V)
            if ctx.zdtype != f32:
                dz = dz.to(ctx.zdtype)
        return dz, g_w, g_b, g_pos, None


class MSELossFn(torch.autograd.Function):
    """forward_loss (tae.py:256-265) with patchify folded into the kernel's indexing."""

    @staticmethod
    def forward(ctx, pred, imgs, p):
        pred_c = pred.contiguous()
        loss, _ = mse_loss(pred_c, imgs, p)
        ctx.p = p
        ctx.save_for_backward(pred_c, imgs)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        pred, imgs = ctx.saved_tensors
        gs = dloss.detach().reshape(1).to(f32).contiguous()
        _, dpred = mse_loss(pred, imgs, ctx.p, want_grad=True, grad_scale=gs)
        return dpred, None, None
