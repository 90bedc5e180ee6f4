"""B200-native Transformer Autoencoder (TAE): drop-in for the model API of the reference's tae.py.

Same public surface as the reference (file:line refer to /root/reference/tae.py):
  * classes PatchEmbed (:29-54), Attention (:56-82), Mlp (:84-105), Block (:107-131), TAE (:133-271) with the same
    constructor signatures, attribute names and parameter registration order, so `state_dict()` has the same
    keys / shapes / order and seeded initialisation consumes the RNG stream identically;
  * `TAE.forward(imgs) -> (loss, pred)` (:267-271), `forward_encoder` (:224-238), `forward_decoder` (:240-254),
    `forward_loss` (:256-265), `patchify` (:196-208), `unpatchify` (:210-222);
  * the 12 zero-argument factories `tae_patch{16,32,64,128}_vocab*_px256` (:434-483).
Additions: `forward(imgs, return_latent=True) -> (loss, pred, latent)`, and `encode` / `decode` aliases.

What differs is everything underneath: the nn.Module tree only *holds* the fp32 master parameters.  Compute runs
through hand-written sm_100a kernels behind the C ABI in include/tae_b200.h (tcgen05/TMEM GEMMs with fused
bias / GELU / residual / pos-embed epilogues, shared-memory attention, fused LayerNorm and loss kernels), driven
by a handful of coarse autograd Functions (one per transformer block) whose backward passes are written by hand.
Numerics follow the reference under `torch.autocast(dtype=bfloat16)`: bf16 GEMM/attention operands with fp32
accumulation, fp32 residual stream, fp32 LayerNorm statistics, fp32 loss.  There is no CPU path.
"""
from __future__ import annotations

import collections.abc
from functools import partial
from itertools import repeat

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_BF16_ROWDOT, EPI_F32_ACC, EPI_F32_RESID


def _pair(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(repeat(x, 2))


# ----------------------------------------------------------------------------------------------------
# Parameter plumbing: bf16 shadows of the fp32 masters, and where parameter gradients are written
# ----------------------------------------------------------------------------------------------------
def shadow_bf16(p: torch.Tensor) -> torch.Tensor:
    """bf16 copy of an fp32 master parameter, refreshed when the master changed (in-place update, .to(), load).

    The fused optimizer (tae_b200.optim.FusedAdamW) keeps the shadow current itself (its kernel writes both).
    """
    s = getattr(p, "_tae_bf16", None)
    if (s is None or getattr(p, "_tae_bf16_version", -1) != p._version
            or getattr(p, "_tae_bf16_ptr", 0) != p.data_ptr() or s.device != p.device):
        src = p.detach()
        if src.dtype != torch.float32:
            src = src.float()
        s = ops.cast_bf16(src.contiguous(), s if (s is not None and s.shape == p.shape and s.device == p.device) else None)
        p._tae_bf16 = s
        p._tae_bf16_version = p._version
        p._tae_bf16_ptr = p.data_ptr()
    return s


def _sink(p):
    """Direct-gradient mode (set up by FusedAdamW): returns (arena view to write into, accumulate flag)."""
    t = getattr(p, "_tae_grad", None)
    if t is None:
        return None, 0
    acc = p._tae_dirty
    g = p.grad
    if g is not t:
        if g is not None and g.data_ptr() != t.data_ptr():
            # autograd's AccumulateGrad got here first (a torch-native op also uses this parameter): keep its gradient
            with torch.no_grad():
                (t.add_ if acc else t.copy_)(g)
            acc = 1
        p.grad = t
    if not p._tae_dirty:
        p._tae_dirty = 1
        p._tae_arena.n_direct += 1
    return t, acc


def _done(p, g):
    """What autograd receives for parameter p: None when the gradient was written in place (direct mode)."""
    if getattr(p, "_tae_grad", None) is not None:
        cb = getattr(p, "_tae_ready", None)
        if cb is not None:
            cb(p)
        return None
    return g


# ---- weight gradients on a side stream (short token grids) -------------------------------------------------------
# With few token rows (patch64: M = 4096, patch128: M = 1024 at B = 256) a dgrad GEMM has fewer output tiles than the GPU
# has SM pairs (40 tiles of 256 x 256 on 74 pairs for the D = 2560 outputs), while the weight-gradient GEMM of the same
# layer — independent of it, same input dy — has hundreds.  The weight gradients are therefore enqueued on a second
# stream: they fill the SMs the short dgrad / LayerNorm / attention kernels of the critical path leave idle.  The side
# stream is forked after dy has been produced and joined at the end of backward (and before a gradient bucket leaves).
WGRAD_OVERLAP_MAX_ROWS = 8192   # 0 disables
_wg_streams: dict = {}          # device index -> side stream
_wg_state = {"join_queued": False, "used": False}


def wgrad_side_stream(device=None, busy_only: bool = False):
    """The side stream carrying weight-gradient GEMMs on `device`, or None if none has been used yet (with `busy_only`:
    None unless it holds work enqueued since the last join — a stream without pending work must not be waited on while a
    CUDA graph is being captured, it is not part of the capture)."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if busy_only and not _wg_state["used"]:
        return None
    return _wg_streams.get(idx)


def join_wgrad_stream():
    """Order the current stream after every weight-gradient GEMM enqueued so far (no host synchronisation)."""
    side = wgrad_side_stream()
    if side is not None and _wg_state["used"]:
        torch.cuda.current_stream().wait_stream(side)
        _wg_state["used"] = False
    _wg_state["join_queued"] = False


def _wgrad(p, dy_b, x_b):
    """dW[out,in] = dy^T x  (fp32, written straight into the gradient arena in direct mode)."""
    t, acc = _sink(p)
    out = None if t is None else t.view(dy_b.shape[1], x_b.shape[1])
    if out is not None and 0 < dy_b.shape[0] <= WGRAD_OVERLAP_MAX_ROWS:
        dev = dy_b.device.index
        side = _wg_streams.get(dev)
        if side is None:
            side = _wg_streams[dev] = torch.cuda.Stream(device=dy_b.device)
            ops.set_dynamic_scheduling(True)  # co-running persistent kernels must draw tiles dynamically to stay balanced
        if not _wg_state["join_queued"]:
            torch.autograd.Variable._execution_engine.queue_callback(join_wgrad_stream)
            _wg_state["join_queued"] = True
        side.wait_stream(torch.cuda.current_stream())  # dy_b / x_b (and earlier writes to `out`) are ordered before
        with torch.cuda.stream(side):
            ops.gemm(dy_b, x_b, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC, out=out, beta=acc)
        dy_b.record_stream(side)  # the caching allocator must not recycle the operands while the side stream reads them
        x_b.record_stream(side)
        _wg_state["used"] = True
        return _done(p, None)
    g = ops.gemm(dy_b, x_b, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC, out=out, beta=acc)
    return _done(p, g.view(p.shape))


def _bgrad(p, dy_b=None, precomputed=None):
    """db = colsum(dy)."""
    t, acc = _sink(p)
    if precomputed is not None:
        if t is None:
            return _done(p, precomputed)
        if acc:
            t.add_(precomputed)
        else:
            t.copy_(precomputed)
        return _done(p, None)
    g = ops.colsum(dy_b, out=t, accumulate=bool(acc))
    return _done(p, g)


# Side channel between consecutive hand-written backward passes: the bf16 copy of a residual-stream gradient and its
# column sums are by-products of the LayerNorm-backward kernel; the consumer (previous block's backward) looks them
# up by the data pointer of the fp32 gradient autograd hands it, and recomputes them if absent.
_SIDE: dict = {}


def _side_put(dres: torch.Tensor, dres_b: torch.Tensor, colsum: torch.Tensor) -> None:
    if len(_SIDE) >= 2:
        _SIDE.clear()
    # the entry keeps `dres` alive, so its address cannot be recycled for another tensor while the entry exists
    _SIDE[dres.data_ptr()] = (dres_b, colsum, dres)


def _side_get(dres: torch.Tensor):
    ent = _SIDE.pop(dres.data_ptr(), None)
    if ent is not None and ent[2].shape == dres.shape and ent[2].dtype == dres.dtype:
        return ent[0], ent[1]
    d2 = dres.reshape(-1, dres.shape[-1])
    dres_b = ops.cast_bf16(d2.contiguous())
    return dres_b, ops.colsum(dres_b)


def _as_f32_2d(g: torch.Tensor, D: int) -> torch.Tensor:
    g = g.reshape(-1, D)
    if g.dtype != torch.float32:
        g = g.float()
    return g.contiguous()


# ----------------------------------------------------------------------------------------------------
# Autograd functions (hand-written backward passes)
# ----------------------------------------------------------------------------------------------------
class _PatchEmbedFn(torch.autograd.Function):
    """PatchEmbed conv (as im2col GEMM) + bias + pos_embed  (tae.py:46-54, :229)."""

    @staticmethod
    def forward(ctx, imgs, w, b, pos, mod):
        p = mod.patch_size[0]
        B = imgs.shape[0]
        N = mod.num_patches
        D = w.shape[0]
        cols = ops.im2col(imgs, p)
        x = ops.gemm(cols, shadow_bf16(w).view(D, -1), epilogue=EPI_F32_RESID, bias=None if b is None else b.detach(),
                     resid=pos.detach().view(N, D), resid_rows=N)
        ctx.mod, ctx.dims = mod, (B, N, D)
        ctx.save_for_backward(cols)
        return x.view(B, N, D)

    @staticmethod
    def backward(ctx, dx):
        (cols,) = ctx.saved_tensors
        B, N, D = ctx.dims
        proj = ctx.mod.proj
        dx = _as_f32_2d(dx, D)
        dx_b, csum = _side_get(dx)
        gw = _wgrad(proj.weight, dx_b, cols) if ctx.needs_input_grad[1] else None
        gb = _bgrad(proj.bias, precomputed=csum) if (proj.bias is not None and ctx.needs_input_grad[2]) else None
        gpos = None
        if ctx.needs_input_grad[3]:
            pos = ctx.mod._pos_param
            t, acc = _sink(pos)
            g = ops.batch_sum(dx, B, N, out=None if t is None else t.view(N, D), accumulate=bool(acc))
            gpos = _done(pos, g.view(1, N, D))
        return None, gw, gb, gpos, None


class _BlockFn(torch.autograd.Function):
    """One pre-LN transformer block  x + attn(norm1(x)); x + mlp(norm2(x))  (tae.py:128-131)."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkvw, qkvb, pw, pb, n2w, n2b, f1w, f1b, f2w, f2b, blk):
        B, N, D = x.shape
        H = blk.attn.num_heads
        hd = D // H
        eps = blk.norm1.eps
        x2 = x.reshape(B * N, D)
        if x2.dtype != torch.float32:
            x2 = x2.float()
        x2 = x2.contiguous()
        det = lambda t: None if t is None else t.detach()
        ln1, mean1, rstd1 = ops.layernorm_fwd(x2, det(n1w), det(n1b), eps)
        qkv = ops.gemm(ln1, shadow_bf16(qkvw), epilogue=EPI_BF16, bias=det(qkvb))
        att, lse = ops.attention_fwd(qkv, B, N, H, hd)
        xm = ops.gemm(att, shadow_bf16(pw), epilogue=EPI_F32_RESID, bias=det(pb), resid=x2)
        ln2, mean2, rstd2 = ops.layernorm_fwd(xm, det(n2w), det(n2b), blk.norm2.eps)
        # gelu'(h), gelu(h); without a backward to come (no_grad: encode.py / evaluate.py) gelu' is not produced
        gp, a = ops.gemm(ln2, shadow_bf16(f1w), epilogue=EPI_BF16_GELU, bias=det(f1b), gelu_grad=any(ctx.needs_input_grad))
        xo = ops.gemm(a, shadow_bf16(f2w), epilogue=EPI_F32_RESID, bias=det(f2b), resid=xm)
        ctx.blk, ctx.dims = blk, (B, N, D, H, hd)
        ctx.save_for_backward(x2, mean1, rstd1, ln1, qkv, att, lse, xm, mean2, rstd2, ln2, gp, a)
        return xo.view(B, N, D)

    @staticmethod
    def backward(ctx, dxo):
        x2, mean1, rstd1, ln1, qkv, att, lse, xm, mean2, rstd2, ln2, gp, a = ctx.saved_tensors
        blk = ctx.blk
        B, N, D, H, hd = ctx.dims
        need = ctx.needs_input_grad
        attn, mlp = blk.attn, blk.mlp
        dres = _as_f32_2d(dxo, D)
        dy_b, csum = _side_get(dres)

        # ---- MLP branch: y = fc2(gelu(fc1(norm2(xm)))) ----
        g_f2b = _bgrad(mlp.fc2.bias, precomputed=csum) if (mlp.fc2.bias is not None and need[12]) else None
        g_f2w = _wgrad(mlp.fc2.weight, dy_b, a) if need[11] else None
        want_b1 = mlp.fc1.bias is not None and need[10]
        # the dgrad epilogue also emits per-32-row column sums of dh: the fc1 bias gradient without re-reading dh
        part = torch.empty(((B * N + 31) // 32, gp.shape[1]), dtype=torch.float32, device=gp.device) if want_b1 else None
        dh = ops.gemm(dy_b, shadow_bf16(mlp.fc2.weight), b_mn=True, epilogue=EPI_BF16_DGELU, aux=gp, colsum_partials=part)
        g_f1b = None
        if want_b1:
            t, acc = _sink(mlp.fc1.bias)
            g_f1b = _done(mlp.fc1.bias, ops.colsum_f32(part, out=t, accumulate=bool(acc)))
        g_f1w = _wgrad(mlp.fc1.weight, dh, ln2) if need[9] else None
        dln2 = ops.gemm(dh, shadow_bf16(mlp.fc1.weight), b_mn=True, epilogue=EPI_BF16)
        del dh
        t_g, a_g = _sink(blk.norm2.weight)
        t_b, a_b = _sink(blk.norm2.bias)
        dres2, dres2_b, dg2, db2, csum2 = ops.layernorm_bwd(dln2, xm, mean2, rstd2, blk.norm2.weight.detach(), dres,
                                                            dgamma=t_g, dbeta=t_b, acc_mask=a_g | (a_b << 1))
        g_n2w, g_n2b = _done(blk.norm2.weight, dg2), _done(blk.norm2.bias, db2)
        del dln2, dres

        # ---- attention branch: y = proj(sdpa(qkv(norm1(x)))) ----
        g_pb = _bgrad(attn.proj.bias, precomputed=csum2) if (attn.proj.bias is not None and need[6]) else None
        g_pw = _wgrad(attn.proj.weight, dres2_b, att) if need[5] else None
        if ops.attention_takes_delta(N, hd):
            # the proj dgrad epilogue also emits delta = rowsum(dO * O) per (image, head, token): the attention kernel
            # then never re-reads O and dO for it
            datt, delta = ops.gemm(dres2_b, shadow_bf16(attn.proj.weight), b_mn=True, epilogue=EPI_BF16_ROWDOT, aux=att,
                                   rowdot_tokens=N)
            dqkv = ops.attention_bwd(qkv, None, datt, lse, B, N, H, hd, delta=delta)
        else:
            datt = ops.gemm(dres2_b, shadow_bf16(attn.proj.weight), b_mn=True, epilogue=EPI_BF16)
            dqkv = ops.attention_bwd(qkv, att, datt, lse, B, N, H, hd)
        del datt
        g_qb = _bgrad(attn.qkv.bias, dy_b=dqkv) if (attn.qkv.bias is not None and need[4]) else None
        g_qw = _wgrad(attn.qkv.weight, dqkv, ln1) if need[3] else None
        dln1 = ops.gemm(dqkv, shadow_bf16(attn.qkv.weight), b_mn=True, epilogue=EPI_BF16)
        del dqkv
        t_g, a_g = _sink(blk.norm1.weight)
        t_b, a_b = _sink(blk.norm1.bias)
        dx, dx_b, dg1, db1, csum1 = ops.layernorm_bwd(dln1, x2, mean1, rstd1, blk.norm1.weight.detach(), dres2,
                                                      dgamma=t_g, dbeta=t_b, acc_mask=a_g | (a_b << 1))
        g_n1w, g_n1b = _done(blk.norm1.weight, dg1), _done(blk.norm1.bias, db1)
        _side_put(dx, dx_b, csum1)
        return (dx.view(B, N, D), g_n1w, g_n1b, g_qw, g_qb, g_pw, g_pb, g_n2w, g_n2b, g_f1w, g_f1b, g_f2w, g_f2b, None)


class _NormLinearFn(torch.autograd.Function):
    """LayerNorm followed by a Linear producing a bf16 tensor: norm -> dict_proj (tae.py:234-237) and
    decoder_norm -> decoder_pred (tae.py:250-253)."""

    @staticmethod
    def forward(ctx, x, nw, nb, w, b, norm, lin):
        B, N, D = x.shape
        x2 = x.reshape(B * N, D)
        if x2.dtype != torch.float32:
            x2 = x2.float()
        x2 = x2.contiguous()
        ln, mean, rstd = ops.layernorm_fwd(x2, nw.detach(), nb.detach(), norm.eps)
        y = ops.gemm(ln, shadow_bf16(w), epilogue=EPI_BF16, bias=None if b is None else b.detach())
        ctx.norm, ctx.lin, ctx.dims = norm, lin, (B, N, D)
        ctx.save_for_backward(x2, mean, rstd, ln)
        return y.view(B, N, -1)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, ln = ctx.saved_tensors
        norm, lin = ctx.norm, ctx.lin
        B, N, D = ctx.dims
        need = ctx.needs_input_grad
        dy_b = dy.reshape(B * N, -1)
        if dy_b.dtype != torch.bfloat16:
            dy_b = dy_b.to(torch.bfloat16)
        dy_b = dy_b.contiguous()
        g_b = _bgrad(lin.bias, dy_b=dy_b) if (lin.bias is not None and need[4]) else None
        g_w = _wgrad(lin.weight, dy_b, ln) if need[3] else None
        dln = ops.gemm(dy_b, shadow_bf16(lin.weight), b_mn=True, epilogue=EPI_BF16)
        t_g, a_g = _sink(norm.weight)
        t_b, a_b = _sink(norm.bias)
        dx, dx_b, dg, db, csum = ops.layernorm_bwd(dln, x2, mean, rstd, norm.weight.detach(), None, dgamma=t_g,
                                                   dbeta=t_b, acc_mask=a_g | (a_b << 1))
        g_nw, g_nb = _done(norm.weight, dg), _done(norm.bias, db)
        _side_put(dx, dx_b, csum)
        return dx.view(B, N, D), g_nw, g_nb, g_w, g_b, None, None


class _EmbedLatentFn(torch.autograd.Function):
    """decoder_embed + decoder_pos_embed  (tae.py:242-245): bf16 latent [B,N,V] -> fp32 residual stream [B,N,D]."""

    @staticmethod
    def forward(ctx, z, w, b, pos, mod):
        B, N, V = z.shape
        D = w.shape[0]
        z2 = z.reshape(B * N, V)
        if z2.dtype != torch.bfloat16:
            z2 = z2.to(torch.bfloat16)
        z2 = z2.contiguous()
        x = ops.gemm(z2, shadow_bf16(w), epilogue=EPI_F32_RESID, bias=None if b is None else b.detach(),
                     resid=pos.detach().view(N, D), resid_rows=N)
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
        dx = _as_f32_2d(dx, D)
        dx_b, csum = _side_get(dx)
        g_b = _bgrad(lin.bias, precomputed=csum) if (lin.bias is not None and need[2]) else None
        g_w = _wgrad(lin.weight, dx_b, z2) if need[1] else None
        g_pos = None
        if need[3]:
            pos = mod.decoder_pos_embed
            t, acc = _sink(pos)
            g = ops.batch_sum(dx, B, N, out=None if t is None else t.view(N, D), accumulate=bool(acc))
            g_pos = _done(pos, g.view(1, N, D))
        dz = None
        if need[0]:
            dz = ops.gemm(dx_b, shadow_bf16(lin.weight), b_mn=True, epilogue=EPI_BF16).view(B, N, V)
            if ctx.zdtype != torch.bfloat16:
                dz = dz.to(ctx.zdtype)
        return dz, g_w, g_b, g_pos, None


class _MSELossFn(torch.autograd.Function):
    """forward_loss (tae.py:256-265) with patchify folded into the kernel's indexing."""

    @staticmethod
    def forward(ctx, pred, imgs, p):
        pred_c = pred.contiguous()
        loss, _ = ops.mse_loss(pred_c, imgs, p, want_grad=False)
        ctx.p = p
        ctx.save_for_backward(pred_c, imgs)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        pred, imgs = ctx.saved_tensors
        gs = dloss.detach().reshape(1).to(torch.float32).contiguous()
        _, dpred = ops.mse_loss(pred, imgs, ctx.p, want_grad=True, grad_scale=gs)
        return dpred, None, None


def _is_fp32(mod) -> bool:
    return getattr(mod, "_tae_precision", "bf16") == "fp32"


def _fp32():
    from . import fp32  # deferred: fp32.py imports _sink/_done from this module

    return fp32


# ----------------------------------------------------------------------------------------------------
# Modules (parameter containers with the reference's names)
# ----------------------------------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    """2D image to patch embedding (tae.py:29-54)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True, bias=True):
        super().__init__()
        img_size, patch_size = _pair(img_size), _pair(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()
        if norm_layer or not flatten or in_chans != 3 or patch_size[0] != patch_size[1] or img_size[0] != img_size[1]:
            raise NotImplementedError("tae_b200.PatchEmbed supports square RGB images, flatten=True, norm_layer=None")
        self._pos_param = None  # set by the owning model: the pos-embed added in the GEMM epilogue

    def forward(self, x, pos_embed=None):
        """Returns the embedded patches WITH `pos_embed` added when given (the add is fused into the GEMM epilogue)."""
        B, C, H, W = x.shape
        assert H == self.img_size[0], f"Input image height ({H}) doesn't match model ({self.img_size[0]})."
        assert W == self.img_size[1], f"Input image width ({W}) doesn't match model ({self.img_size[1]})."
        _lib.require_device()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        fn = _fp32().PatchEmbedFn if _is_fp32(self) else _PatchEmbedFn
        if pos_embed is None:
            pos_embed = torch.zeros(1, self.num_patches, self.proj.weight.shape[0], device=x.device)
            return fn.apply(x, self.proj.weight, self.proj.bias, pos_embed, self)
        object.__setattr__(self, "_pos_param", pos_embed)
        return fn.apply(x, self.proj.weight, self.proj.bias, pos_embed, self)


class Attention(nn.Module):
    """Scaled dot-product attention (tae.py:56-82); parameter container — compute lives in Block."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, norm_layer=nn.LayerNorm):
        super().__init__()
        assert dim % num_heads == 0, 'dim should be divisible by num_heads'
        if qk_norm:
            raise NotImplementedError("qk_norm is not used by any TAE model and is not implemented")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.proj = nn.Linear(dim, dim)


class Mlp(nn.Module):
    """Transformer MLP (tae.py:84-105); parameter container — compute lives in Block."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None, bias=True,
                 use_conv=False):
        super().__init__()
        if use_conv or norm_layer is not None or act_layer is not nn.GELU:
            raise NotImplementedError("tae_b200.Mlp implements Linear -> exact GELU -> Linear only")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        bias = _pair(bias)
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0])
        self.act = act_layer()
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1])


class Block(nn.Module):
    """Pre-LN transformer block (tae.py:107-131), executed as ONE autograd node of fused kernels."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_norm=False, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, mlp_layer=Mlp):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm, norm_layer=norm_layer)
        self.norm2 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)
        if not isinstance(self.norm1, nn.LayerNorm):
            raise NotImplementedError("tae_b200.Block needs nn.LayerNorm as norm_layer")

    def forward(self, x):
        _lib.require_device()
        a, m = self.attn, self.mlp
        fn = _fp32().BlockFn if _is_fp32(self) else _BlockFn
        return fn.apply(x, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                              self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, self)


class TAE(nn.Module):
    """Transformer-based autoencoder with ViT encoder / decoder (tae.py:133-271)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, vocab_size=16, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4., norm_layer=nn.LayerNorm):
        super().__init__()
        # encoder
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias=True, norm_layer=norm_layer)
                                     for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.dict_proj = nn.Linear(embed_dim, vocab_size, bias=False)
        # decoder
        self.decoder_embed = nn.Linear(vocab_size, decoder_embed_dim, bias=True)
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches, decoder_embed_dim))
        self.decoder_blocks = nn.ModuleList([Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, qkv_bias=True,
                                                   norm_layer=norm_layer) for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.decoder_pred = nn.Linear(decoder_embed_dim, patch_size ** 2 * in_chans, bias=True)
        self.initialize_weights()

    # -- initialisation: same distributions in the same order as tae.py:174-194 --
    def initialize_weights(self):
        torch.nn.init.trunc_normal_(self.pos_embed, std=0.02)
        torch.nn.init.trunc_normal_(self.decoder_pos_embed, std=0.02)
        w = self.patch_embed.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def invalidate_shadows(self):
        """Drop the cached bf16 weight copies (needed only after writing parameters through `.data`)."""
        for p in self.parameters():
            if hasattr(p, "_tae_bf16_version"):
                p._tae_bf16_version = -1

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_shadows()
        return out

    def set_precision(self, precision: str):
        """"bf16" (default): the rounding points of `torch.autocast(dtype=bfloat16)` around the reference — the fast path.
        "fp32": the reference without autocast; every activation fp32, GEMMs as 3-way bf16 splits on the tensor cores
        (tae_b200/fp32.py).  Returns self."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        for m in self.modules():
            object.__setattr__(m, "_tae_precision", precision)
        return self

    @property
    def precision(self) -> str:
        return getattr(self, "_tae_precision", "bf16")

    # -- integer patch index maps (bit-exact) --
    def patchify(self, imgs):
        """imgs (N, 3, H, W) -> (N, L, patch_size**2 * 3)   (tae.py:196-208)"""
        p = self.patch_embed.patch_size[0]
        assert imgs.shape[2] == imgs.shape[3] and imgs.shape[2] % p == 0
        return ops.patchify(imgs, p)

    def unpatchify(self, x):
        """x (N, L, patch_size**2 * 3) -> (N, 3, H, W)   (tae.py:210-222)"""
        p = self.patch_embed.patch_size[0]
        h = int(x.shape[1] ** .5)
        assert h * h == x.shape[1]
        return ops.unpatchify(x, p)

    # -- forward passes --
    def forward_encoder(self, x):
        x = self.patch_embed(x, self.pos_embed)  # conv-as-GEMM, bias and pos_embed fused in the epilogue
        for blk in self.blocks:
            x = blk(x)
        fn = _fp32().NormLinearFn if _is_fp32(self) else _NormLinearFn
        return fn.apply(x, self.norm.weight, self.norm.bias, self.dict_proj.weight, self.dict_proj.bias,
                        self.norm, self.dict_proj)

    def forward_decoder(self, x):
        _lib.require_device()
        fp = _is_fp32(self)
        fn = _fp32().EmbedLatentFn if fp else _EmbedLatentFn
        x = fn.apply(x, self.decoder_embed.weight, self.decoder_embed.bias, self.decoder_pos_embed, self)
        for blk in self.decoder_blocks:
            x = blk(x)
        fn = _fp32().NormLinearFn if fp else _NormLinearFn
        return fn.apply(x, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_pred.weight,
                        self.decoder_pred.bias, self.decoder_norm, self.decoder_pred)

    def forward_loss(self, imgs, pred):
        """imgs [N, 3, H, W], pred [N, L, p*p*3] -> mean squared error per pixel (fp32 scalar)."""
        _lib.require_device()
        if imgs.dtype != torch.float32:
            imgs = imgs.float()
        if _is_fp32(self):
            if pred.dtype != torch.float32:
                pred = pred.float()
            return _fp32().MSELossFn.apply(pred, imgs.contiguous(), self.patch_embed.patch_size[0])
        if pred.dtype != torch.bfloat16:
            pred = pred.to(torch.bfloat16)
        return _MSELossFn.apply(pred, imgs.contiguous(), self.patch_embed.patch_size[0])

    def forward(self, imgs, return_latent=False):
        latent = self.forward_encoder(imgs)
        pred = self.forward_decoder(latent)
        loss = self.forward_loss(imgs, pred)
        if return_latent:
            return loss, pred, latent
        return loss, pred

    # aliases named by the north-star API
    encode = forward_encoder
    decode = forward_decoder


# ----------------------------------------------------------------------------------------------------
# Downstream consumers of the same Block: VITForRecognition (tae.py:274-342), VITForSegmentation (tae.py:345-429)
# (SURVEY.md §8f.2).  They reuse the TAE decoder's autograd functions unchanged; the new pieces are a standalone
# LayerNorm, mean-pool + head, and the C-channel unpatchify.
# ----------------------------------------------------------------------------------------------------
class _NormFn(torch.autograd.Function):
    """decoder_norm on its own (tae.py:329): fp32 in, fp32 out (LayerNorm stays fp32 under autocast)."""

    @staticmethod
    def forward(ctx, x, w, b, norm):
        from . import fp32 as F

        B, N, D = x.shape
        x2 = _as_f32_2d(x, D)
        y, mean, rstd = F.layernorm_fwd(x2, w.detach(), b.detach(), norm.eps)
        ctx.norm, ctx.dims = norm, (B, N, D)
        ctx.save_for_backward(x2, mean, rstd)
        return y.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        from . import fp32 as F

        x2, mean, rstd = ctx.saved_tensors
        norm = ctx.norm
        B, N, D = ctx.dims
        t_g, a_g = _sink(norm.weight)
        t_b, a_b = _sink(norm.bias)
        dx, dg, db = F.layernorm_bwd(_as_f32_2d(dy, D), x2, mean, rstd, norm.weight.detach(), None, t_g, t_b,
                                     a_g | (a_b << 1))
        return dx.view(B, N, D), _done(norm.weight, dg), _done(norm.bias, db), None


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class _MeanHeadFn(torch.autograd.Function):
    """forward_head (tae.py:332-335): global average pool over tokens (fp32) -> head Linear (bf16 out under autocast).
    Class counts and batch sizes that are not multiples of 8 are zero-padded to the GEMM's granularity."""

    @staticmethod
    def forward(ctx, x, w, b, head, fp32_mode):
        B, N, D = x.shape
        pooled = ops.token_mean(_as_f32_2d(x, D), B, N)  # fp32 [B, D]
        C = w.shape[0]
        Bp, Cp = _pad8(B), _pad8(C)
        if fp32_mode:
            from . import fp32 as F

            pp = pooled if Bp == B else torch.cat([pooled, pooled.new_zeros(Bp - B, D)])
            wp = w.detach() if Cp == C else torch.cat([w.detach(), w.new_zeros(Cp - C, D)])
            y = F.gemm_f32(F.split3(pp), F.split3(wp))
            if b is not None:
                bp = b.detach() if Cp == C else torch.cat([b.detach(), b.new_zeros(Cp - C)])
                F.bias_act(y, bp)
        else:
            pp = torch.zeros((Bp, D), dtype=torch.bfloat16, device=x.device)
            ops.cast_bf16(pooled, pp[:B])
            wb = shadow_bf16(w)
            if Cp != C:
                wb = torch.cat([wb, wb.new_zeros(Cp - C, D)])
            bp = None
            if b is not None:
                bp = b.detach() if Cp == C else torch.cat([b.detach(), b.new_zeros(Cp - C)])
            y = ops.gemm(pp, wb, epilogue=EPI_BF16, bias=bp)
        ctx.head, ctx.dims, ctx.fp32_mode = head, (B, N, D, C, Bp, Cp), fp32_mode
        ctx.save_for_backward(pp)
        return y[:B, :C]

    @staticmethod
    def backward(ctx, dy):
        (pp,) = ctx.saved_tensors
        head = ctx.head
        B, N, D, C, Bp, Cp = ctx.dims
        need = ctx.needs_input_grad
        w, b = head.weight, head.bias
        dyp = torch.zeros((Bp, Cp), dtype=torch.float32 if ctx.fp32_mode else torch.bfloat16, device=dy.device)
        dyp[:B, :C] = dy
        g_b = None
        if b is not None and need[2]:
            t, acc = _sink(b)
            full = ops.colsum_f32(dyp) if ctx.fp32_mode else ops.colsum(dyp)
            if t is None:
                g_b = full[:C].clone()
            else:
                t.add_(full[:C]) if acc else t.copy_(full[:C])
            g_b = _done(b, g_b)
        g_w = None
        if ctx.fp32_mode:
            from . import fp32 as F

            dy3 = F.split3(dyp)
            gw_full = F.gemm_f32(dy3, F.split3(pp), a_mn=True, b_mn=True) if need[1] else None
            wp = w.detach() if Cp == C else torch.cat([w.detach(), w.new_zeros(Cp - C, D)])
            dpooled = F.gemm_f32(dy3, F.split3(wp), b_mn=True)[:B] if need[0] else None
        else:
            gw_full = ops.gemm(dyp, pp, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC, splits=1) if need[1] else None
            wb = shadow_bf16(w)
            if Cp != C:
                wb = torch.cat([wb, wb.new_zeros(Cp - C, D)])
            dpooled = ops.gemm(dyp, wb, b_mn=True, epilogue=EPI_F32_ACC, splits=1)[:B] if need[0] else None
        if need[1]:
            t, acc = _sink(w)
            if t is None:
                g_w = gw_full[:C].contiguous()
            else:
                tv = t.view(C, D)
                tv.add_(gw_full[:C]) if acc else tv.copy_(gw_full[:C])
            g_w = _done(w, g_w)
        dx = ops.token_mean_bwd(dpooled.contiguous(), B, N).view(B, N, D) if need[0] else None
        return dx, g_w, g_b, None, None


class _UnpatchifyCFn(torch.autograd.Function):
    """VITForSegmentation.unpatchify (tae.py:391-403): a pure permutation, so its adjoint is the inverse permutation."""

    @staticmethod
    def forward(ctx, x, p, C):
        ctx.p = p
        return ops.unpatchify_c(x, p, C)

    @staticmethod
    def backward(ctx, dimgs):
        return ops.patchify_c(dimgs, ctx.p), None, None


class _ViTBase(nn.Module):
    def initialize_weights(self):
        torch.nn.init.trunc_normal_(self.decoder_pos_embed, std=0.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    set_precision = TAE.set_precision
    precision = TAE.precision
    invalidate_shadows = TAE.invalidate_shadows

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_shadows()
        return out

    def _embed(self, x):
        _lib.require_device()
        fn = _fp32().EmbedLatentFn if _is_fp32(self) else _EmbedLatentFn
        return fn.apply(x, self.decoder_embed.weight, self.decoder_embed.bias, self.decoder_pos_embed, self)


class VITForRecognition(_ViTBase):
    """PatchEmbed-less ViT classifier over TAE latents (tae.py:274-342): same constructor, attributes and state_dict."""

    def __init__(self, num_patches=256, vocab_size=16, decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16,
                 mlp_ratio=4., norm_layer=nn.LayerNorm, num_classes=None):
        super().__init__()
        self.decoder_embed = nn.Linear(vocab_size, decoder_embed_dim, bias=True)
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches, decoder_embed_dim))
        self.decoder_blocks = nn.ModuleList([Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, qkv_bias=True,
                                                   norm_layer=norm_layer) for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.head = nn.Linear(decoder_embed_dim, num_classes, bias=True) if num_classes is not None else nn.Identity()
        self.initialize_weights()

    def forward_features(self, x):
        x = self._embed(x)
        for blk in self.decoder_blocks:
            x = blk(x)
        return _NormFn.apply(x, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_norm)

    def forward_head(self, x):
        if isinstance(self.head, nn.Identity):
            B, N, D = x.shape
            return _MeanOnlyFn.apply(x)
        return _MeanHeadFn.apply(x, self.head.weight, self.head.bias, self.head, _is_fp32(self))

    def forward(self, x):
        return self.forward_head(self.forward_features(x))


class _MeanOnlyFn(torch.autograd.Function):
    """x.mean(dim=1) when the classifier has no head (num_classes=None)."""

    @staticmethod
    def forward(ctx, x):
        B, N, D = x.shape
        ctx.dims = (B, N, D)
        return ops.token_mean(_as_f32_2d(x, D), B, N)

    @staticmethod
    def backward(ctx, dy):
        B, N, D = ctx.dims
        return ops.token_mean_bwd(_as_f32_2d(dy, D), B, N).view(B, N, D)


class VITForSegmentation(_ViTBase):
    """PatchEmbed-less ViT segmenter over TAE latents with an auxiliary head at 3/4 depth (tae.py:345-429)."""

    def __init__(self, num_patches=256, patch_size=16, vocab_size=16, decoder_embed_dim=512, decoder_depth=8,
                 decoder_num_heads=16, mlp_ratio=4., norm_layer=nn.LayerNorm, num_classes=None):
        super().__init__()
        self.aux_depth = int(decoder_depth * 0.75)
        self.patch_size = patch_size
        self.num_classes = num_classes
        self.decoder_embed = nn.Linear(vocab_size, decoder_embed_dim, bias=True)
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches, decoder_embed_dim))
        self.decoder_blocks = nn.ModuleList([Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, qkv_bias=True,
                                                   norm_layer=norm_layer) for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.aux_decoder_norm = norm_layer(decoder_embed_dim)
        self.head = nn.Linear(decoder_embed_dim, patch_size ** 2 * num_classes, bias=True)
        self.aux_head = nn.Linear(decoder_embed_dim, patch_size ** 2 * num_classes, bias=True)
        self.initialize_weights()

    def unpatchify(self, x):
        """x (N, L, patch_size**2 * C) -> (N, C, H, W)   (tae.py:391-403)"""
        h = int(x.shape[1] ** .5)
        assert h * h == x.shape[1]
        return _UnpatchifyCFn.apply(x, self.patch_size, x.shape[2] // (self.patch_size ** 2))

    def _norm_head(self, x, norm, lin):
        fn = _fp32().NormLinearFn if _is_fp32(self) else _NormLinearFn
        return fn.apply(x, norm.weight, norm.bias, lin.weight, lin.bias, norm, lin)

    def forward(self, x):
        x = self._embed(x)
        result = collections.OrderedDict()
        aux = None
        for i, blk in enumerate(self.decoder_blocks):
            x = blk(x)
            if i + 1 == self.aux_depth:
                aux = self.unpatchify(self._norm_head(x, self.aux_decoder_norm, self.aux_head))
        x = self.unpatchify(self._norm_head(x, self.decoder_norm, self.head))
        result["out"] = x
        result["aux"] = aux
        return result


# ----------------------------------------------------------------------------------------------------
# Model zoo (tae.py:431-483): same names, zero arguments
# ----------------------------------------------------------------------------------------------------
_ZOO = {
    # patch: (embed_dim, depth, heads, vocab sizes)
    16: (1024, 15, 16, (16, 64, 256)),
    32: (2048, 18, 32, (64, 256, 1024)),
    64: (2560, 21, 32, (256, 1024, 4096)),
    128: (2560, 22, 32, (1024, 4096, 16384)),
}


def _make_factory(patch, vocab, dim, depth, heads):
    def factory():
        return TAE(patch_size=patch, vocab_size=vocab, img_size=256, embed_dim=dim, depth=depth, num_heads=heads,
                   decoder_embed_dim=dim, decoder_depth=depth, decoder_num_heads=heads, mlp_ratio=4,
                   norm_layer=partial(nn.LayerNorm, eps=1e-6))

    factory.__name__ = f"tae_patch{patch}_vocab{vocab}_px256"
    factory.__qualname__ = factory.__name__
    factory.__doc__ = f"TAE with {patch}x{patch} patches, latent width {vocab}, 256px inputs (reference tae.py:434-483)."
    return factory


MODEL_NAMES = []
for _p, (_dim, _depth, _heads, _vocabs) in _ZOO.items():
    for _v in _vocabs:
        _f = _make_factory(_p, _v, _dim, _depth, _heads)
        globals()[_f.__name__] = _f
        MODEL_NAMES.append(_f.__name__)
del _p, _dim, _depth, _heads, _vocabs, _v, _f


# ************** RECOGNITION / SEGMENTATION (tae.py:485-587): ViT-Base over TAE latents **************
_VIT_ZOO = {256: (16, 64, 256), 64: (64, 256, 1024), 16: (256, 1024, 4096), 4: (1024, 4096, 16384)}


def _make_vit_factory(kind, cls, num_patches, vocab):
    def factory(num_classes=None):
        return cls(num_patches=num_patches, vocab_size=vocab, decoder_embed_dim=768, decoder_depth=12, decoder_num_heads=12,
                   mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), num_classes=num_classes)

    factory.__name__ = f"vit_{kind}_numpatches{num_patches}_vocab{vocab}_base"
    factory.__qualname__ = factory.__name__
    factory.__doc__ = f"ViT-Base {kind} model over {num_patches}-token, width-{vocab} TAE latents (reference tae.py:485-587)."
    return factory


VIT_MODEL_NAMES = []
for _kind, _cls in (("recognition", VITForRecognition), ("segmentation", VITForSegmentation)):
    for _np, _vocabs in _VIT_ZOO.items():
        for _v in _vocabs:
            _f = _make_vit_factory(_kind, _cls, _np, _v)
            globals()[_f.__name__] = _f
            VIT_MODEL_NAMES.append(_f.__name__)
del _kind, _cls, _np, _vocabs, _v, _f
