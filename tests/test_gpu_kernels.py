"""Per-kernel parity tests: every C-ABI entry point against the oracle's restatement of the same op (GPU only).

Tolerances: bf16 outputs are compared within a few bf16 ulps of the result magnitude (2e-2 relative is the
north-star gate for bf16); fp32 outputs of fp32 inputs within 1e-4 relative; integer index maps bit-exact.
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tae_oracle as O  # noqa: E402  (checker only)


def _ops():
    from tae_b200 import ops

    return ops


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_err_scaled(a, b) -> float:
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def randn(*shape, dtype=torch.bfloat16, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    (128, 256, 64), (256, 512, 128), (512, 768, 256), (384, 1024, 1024), (1000, 264, 72), (4, 16, 128), (128, 16, 1024),
    (1024, 1024, 16), (130, 3072, 1024), (2048, 4096, 1024),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_majors(M, N, K, a_mn, b_mn):
    ops = _ops()
    from tae_b200._lib import EPI_BF16

    if a_mn and M % 8:
        pytest.skip("MN-major A needs M % 8 == 0")
    A = randn(M, K, seed=1)
    B = randn(N, K, seed=2)
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    out = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, epilogue=EPI_BF16)
    torch.cuda.synchronize()
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    assert max_err_scaled(out.float(), ref) < 1e-2, (rel_err(out.float(), ref))
    assert rel_err(out.float(), ref) < 5e-3


@pytest.mark.parametrize("M,N,K", [(640, 512, 256), (1000, 264, 72), (300, 4096, 64), (2048, 1032, 128), (96, 256, 64)])
def test_gemm_epilogue_bias_gelu(M, N, K):
    """fc1 + GELU epilogue, including ragged edges: N % 32 != 0 (a partly out-of-range 32-column step), M % 32 != 0,
    epilogue warps whose 32 rows lie entirely outside the matrix, and the single-CTA kernel (M <= 128)."""
    ops = _ops()
    from tae_b200._lib import EPI_BF16, EPI_BF16_GELU

    A, B = randn(M, K, seed=3), randn(N, K, seed=4, scale=0.1)
    bias = randn(N, dtype=torch.float32, seed=5)
    acc = A.float() @ B.float().t() + bias.to(torch.bfloat16).float()
    out = ops.gemm(A, B, epilogue=EPI_BF16, bias=bias)
    assert rel_err(out.float(), acc) < 5e-3
    gp, a = ops.gemm(A, B, epilogue=EPI_BF16_GELU, bias=bias)  # out = gelu'(h), out2 = gelu(h), h = bf16(acc + bias)
    hf = acc.to(torch.bfloat16).float()
    aref = O.gelu(hf)
    gpref = 0.5 * (1 + torch.erf(hf / math.sqrt(2))) + hf * torch.exp(-0.5 * hf * hf) / math.sqrt(2 * math.pi)
    assert max_err_scaled(a.float(), aref) < 1e-2 and rel_err(a.float(), aref) < 4e-3
    assert max_err_scaled(gp.float(), gpref) < 1e-2 and rel_err(gp.float(), gpref) < 4e-3
    # inference form (out == NULL): the same gelu(h), no gelu' written
    none, a_inf = ops.gemm(A, B, epilogue=EPI_BF16_GELU, bias=bias, gelu_grad=False)
    assert none is None and torch.equal(a_inf, a)
    # the two outputs are views into larger buffers in the model (leading dimension > N): nothing outside [M, N] is touched
    big = torch.full((M + 8, N + 64), 7.0, dtype=torch.bfloat16, device="cuda")
    big2 = torch.full((M + 8, N + 64), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, B, epilogue=EPI_BF16_GELU, bias=bias, out=big[:M, :N], out2=big2[:M, :N])
    assert torch.equal(big[:M, :N], gp) and torch.equal(big2[:M, :N], a)
    assert bool((big[M:] == 7).all()) and bool((big[:, N:] == 7).all())
    assert bool((big2[M:] == 7).all()) and bool((big2[:, N:] == 7).all())


@pytest.mark.parametrize("M,N,K", [(1024, 2560, 2560), (1024, 7680, 512), (1024, 10240, 256), (4096, 2560, 640), (4096, 4096, 256),
                                   (1024, 2568, 320), (520, 1160, 128), (2048, 16384, 64)])
def test_gemm_narrow_tiles_for_short_grids(M, N, K):
    """K-major-B GEMMs with few row tiles pick a tile width below 256 against wave quantisation (160 / 192 / 224 / 128 for
    these shapes; the last three also have ragged N or M): every forward epilogue must be unaffected by the width."""
    ops = _ops()
    from tae_b200._lib import EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32_RESID

    A, B = randn(M, K, seed=80), randn(N, K, seed=81, scale=0.1)
    bias = randn(N, dtype=torch.float32, seed=82)
    acc = A.float() @ B.float().t() + bias.to(torch.bfloat16).float()
    big = torch.full((M + 2, N + 40), 7.0, dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(A, B, epilogue=EPI_BF16, bias=bias, out=big[:M, :N])
    assert rel_err(out.float(), acc) < 5e-3 and max_err_scaled(out.float(), acc) < 1e-2
    assert bool((big[M:] == 7).all()) and bool((big[:, N:] == 7).all())
    gp, a = ops.gemm(A, B, epilogue=EPI_BF16_GELU, bias=bias)
    hf = acc.to(torch.bfloat16).float()
    assert rel_err(a.float(), O.gelu(hf)) < 4e-3
    gpref = 0.5 * (1 + torch.erf(hf / math.sqrt(2))) + hf * torch.exp(-0.5 * hf * hf) / math.sqrt(2 * math.pi)
    assert rel_err(gp.float(), gpref) < 4e-3
    resid = randn(M, N, dtype=torch.float32, seed=83)
    y = acc.to(torch.bfloat16).float()
    outr = ops.gemm(A, B, epilogue=EPI_F32_RESID, bias=bias, resid=resid)
    assert rel_err(outr - resid, y) < 5e-3
    if M % 8 == 0:
        pos = randn(8, N, dtype=torch.float32, seed=84)
        outp = ops.gemm(A, B, epilogue=EPI_F32_RESID, bias=bias, resid=pos, resid_rows=8)
        assert rel_err(outp - pos.repeat(M // 8, 1), y) < 5e-3
    gpm = randn(M, N, seed=85, scale=0.5)
    part = torch.empty((M + 31) // 32, N, device="cuda")
    dh = ops.gemm(A, B, epilogue=EPI_BF16_DGELU, aux=gpm, colsum_partials=part)
    ref = (A.float() @ B.float().t()).to(torch.bfloat16).float() * gpm.float()
    assert rel_err(dh.float(), ref) < 5e-3
    assert rel_err(ops.colsum_f32(part), dh.float().sum(0)) < 1e-4


def test_gemm_epilogue_residual_and_posembed():
    ops = _ops()
    from tae_b200._lib import EPI_F32_RESID

    M, N, K = 512, 256, 128
    A, B = randn(M, K, seed=6), randn(N, K, seed=7, scale=0.2)
    bias = randn(N, dtype=torch.float32, seed=8)
    resid = randn(M, N, dtype=torch.float32, seed=9)
    y = (A.float() @ B.float().t() + bias.to(torch.bfloat16).float()).to(torch.bfloat16).float()
    out = ops.gemm(A, B, epilogue=EPI_F32_RESID, bias=bias, resid=resid)
    assert out.dtype == torch.float32
    assert max_err_scaled(out - resid, y) < 1e-2
    # broadcast residual (pos-embed): rows repeat every 64 tokens
    pos = randn(64, N, dtype=torch.float32, seed=10)
    out2 = ops.gemm(A, B, epilogue=EPI_F32_RESID, bias=bias, resid=pos, resid_rows=64)
    ref2 = y + pos.repeat(M // 64, 1)
    assert max_err_scaled(out2, ref2) < 1e-2


@pytest.mark.parametrize("M,N,K", [(1000, 264, 72), (300, 1032, 64), (2048, 1024, 1024), (4096, 1024, 4096)])
@pytest.mark.parametrize("inplace", [False, True])
def test_gemm_epilogue_residual_ragged(M, N, K, inplace):
    """fp32 residual epilogue at ragged edges (N % 16 != 0, M % 32 != 0), strided resid / out views and out aliasing resid."""
    ops = _ops()
    from tae_b200._lib import EPI_F32_RESID

    A, B = randn(M, K, seed=70), randn(N, K, seed=71, scale=0.2)
    bias = randn(N, dtype=torch.float32, seed=72)
    big = torch.full((M + 3, N + 20), 7.0, dtype=torch.float32, device="cuda")
    resid_big = randn(M, N + 12, dtype=torch.float32, seed=73)
    y = (A.float() @ B.float().t() + bias.to(torch.bfloat16).float()).to(torch.bfloat16).float()
    if inplace:
        big[:M, :N] = resid_big[:, :N]
        want = big[:M, :N] + y
        ops.gemm(A, B, epilogue=EPI_F32_RESID, bias=bias, resid=big[:M, :N], out=big[:M, :N])
    else:
        want = resid_big[:, :N] + y
        ops.gemm(A, B, epilogue=EPI_F32_RESID, bias=bias, resid=resid_big[:, :N], out=big[:M, :N])
    assert max_err_scaled(big[:M, :N] - want + y, y) < 1e-2 and rel_err(big[:M, :N], want) < 2e-3
    assert bool((big[M:] == 7).all()) and bool((big[:, N:] == 7).all())


@pytest.mark.parametrize("M,N,K,tokens,with_bias", [(4096, 1024, 1024, 256, False), (1000, 192, 72, 100, True),
                                                     (300, 64, 264, 4, True), (2560, 2560, 128, 16, False)])
def test_gemm_rowdot_epilogue_ragged(M, N, K, tokens, with_bias):
    ops = _ops()
    A, W = randn(M, K, seed=74, scale=0.5), randn(K, N, seed=75, scale=0.5)
    aux_big = randn(M, N + 8, seed=76)
    aux = aux_big[:, :N]
    bias = randn(N, dtype=torch.float32, seed=77) if with_bias else None
    out_big = torch.full((M + 2, N + 16), 7.0, dtype=torch.bfloat16, device="cuda")
    out, rd = ops.gemm(A, W, b_mn=True, epilogue=5, aux=aux, rowdot_tokens=tokens, bias=bias, out=out_big[:M, :N])
    ref = A.float() @ W.float() + (bias.to(torch.bfloat16).float() if with_bias else 0)
    assert max_err_scaled(out.float(), ref) < 1e-2
    want = (out.float() * aux.float()).view(M // tokens, tokens, N // 64, 64).sum(-1).permute(0, 2, 1).contiguous()
    assert rd.shape == (M // tokens, N // 64, tokens) and rel_err(rd, want) < 1e-5
    assert bool((out_big[M:] == 7).all()) and bool((out_big[:, N:] == 7).all())


@pytest.mark.parametrize("splits", [1, 3, 0])
@pytest.mark.parametrize("beta", [0, 1])
def test_gemm_wgrad_accumulate(splits, beta):
    ops = _ops()
    from tae_b200._lib import EPI_F32_ACC

    T, Nout, Kin = 4096, 384, 256  # dW[Nout,Kin] = dY[T,Nout]^T X[T,Kin]
    dY, X = randn(T, Nout, seed=11, scale=0.1), randn(T, Kin, seed=12)
    ref = dY.float().t() @ X.float()
    out = randn(Nout, Kin, dtype=torch.float32, seed=13)
    prev = out.clone()
    ops.gemm(dY, X, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC, out=out, beta=beta, splits=splits)
    want = ref + (prev if beta else 0)
    assert rel_err(out, want) < 1e-3


def test_gemm_dgelu_epilogue():
    ops = _ops()
    from tae_b200._lib import EPI_BF16_DGELU

    M, N, K = 384, 512, 128  # da[M,N] = dy[M,K] W[K,N];  dh = bf16(da) * gp,  gp = gelu'(h) saved by the forward
    dy, W = randn(M, K, seed=14), randn(K, N, seed=15, scale=0.1)
    gp = randn(M, N, seed=16, scale=0.5)
    part = torch.empty((M + 31) // 32, N, device="cuda")
    out = ops.gemm(dy, W, b_mn=True, epilogue=EPI_BF16_DGELU, aux=gp, colsum_partials=part)
    da = (dy.float() @ W.float()).to(torch.bfloat16).float()
    ref = da * gp.float()
    assert max_err_scaled(out.float(), ref) < 1e-2
    assert rel_err(out.float(), ref) < 5e-3
    # the epilogue's column-sum by-product == sums of the bf16 output it wrote (32-row groups), and its reduction
    want = out.float().reshape(M // 32, 32, N).sum(1)
    assert rel_err(part, want) < 1e-5
    assert rel_err(ops.colsum_f32(part), out.float().sum(0)) < 1e-5
    # ragged M: rows beyond M contribute nothing
    M2 = 200
    part2 = torch.full(((M2 + 31) // 32, N), 7.0, device="cuda")
    out2 = ops.gemm(dy[:M2].contiguous(), W, b_mn=True, epilogue=EPI_BF16_DGELU, aux=gp[:M2].contiguous(), colsum_partials=part2)
    assert rel_err(ops.colsum_f32(part2), out2.float().sum(0)) < 1e-5


@pytest.mark.parametrize("M,N,K,with_partials", [(1000, 264, 72, True), (2048, 4096, 1024, True), (300, 1032, 64, True),
                                                  (520, 768, 256, False)])
def test_gemm_dgelu_epilogue_ragged(M, N, K, with_partials):
    """GELU' epilogue at ragged edges: N % 32 != 0, M % 32 != 0, warps whose rows lie outside the matrix, strided aux and
    out (views into wider buffers), with and without the column-sum by-product."""
    ops = _ops()
    from tae_b200._lib import EPI_BF16_DGELU

    dy, W = randn(M, K, seed=60), randn(K, N, seed=61, scale=0.1)
    gp_big = randn(M, N + 40, seed=62, scale=0.5)
    gp = gp_big[:, :N]
    out_big = torch.full((M + 4, N + 24), 7.0, dtype=torch.bfloat16, device="cuda")
    nparts = (M + 31) // 32
    part = torch.full((nparts, N), 7.0, device="cuda") if with_partials else None
    out = ops.gemm(dy, W, b_mn=True, epilogue=EPI_BF16_DGELU, aux=gp, out=out_big[:M, :N], colsum_partials=part)
    da = (dy.float() @ W.float()).to(torch.bfloat16).float()
    ref = da * gp.float()
    assert max_err_scaled(out.float(), ref) < 1e-2 and rel_err(out.float(), ref) < 5e-3
    assert bool((out_big[M:] == 7).all()) and bool((out_big[:, N:] == 7).all())
    if with_partials:
        pad = torch.zeros(nparts * 32, N, device="cuda")
        pad[:M] = out.float()
        assert rel_err(part, pad.view(nparts, 32, N).sum(1)) < 1e-5


def test_gemm_rejects_bad_shapes():
    ops = _ops()
    from tae_b200._lib import TaeError

    with pytest.raises(TaeError):
        ops.gemm(randn(8, 12), randn(16, 12))  # K % 8 != 0
    with pytest.raises(TaeError):
        ops.gemm(torch.zeros(8, 16, dtype=torch.bfloat16), torch.zeros(16, 16, dtype=torch.bfloat16))  # CPU tensors


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,D", [(512, 1024), (1030, 2048), (77, 2560), (256, 128), (64, 640)])
def test_layernorm_fwd_bwd(rows, D):
    ops = _ops()
    x = randn(rows, D, dtype=torch.float32, seed=20) * 2 + 0.5
    w = randn(D, dtype=torch.float32, seed=21) * 0.2 + 1
    b = randn(D, dtype=torch.float32, seed=22) * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
    xr = x.detach().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = O.layer_norm(xr, wr, br, 1e-6)
    assert max_err_scaled(y.float(), yr.detach()) < 8e-3
    assert rel_err(mean, x.mean(-1)) < 1e-5
    assert rel_err(rstd, 1 / torch.sqrt(x.var(-1, unbiased=False) + 1e-6)) < 1e-5
    dy = randn(rows, D, seed=23)
    dres = randn(rows, D, dtype=torch.float32, seed=24)
    yr.backward(dy.float())
    dx, dx_b, dg, db, cs = ops.layernorm_bwd(dy, x, mean, rstd, w, dres)
    assert rel_err(dx - dres, xr.grad) < 1e-4
    assert rel_err(dg, wr.grad) < 1e-4
    assert rel_err(db, br.grad) < 1e-4
    assert torch.equal(dx_b, dx.to(torch.bfloat16))
    assert rel_err(cs, dx_b.float().sum(0)) < 1e-4
    # accumulate mask: dgamma += , dbeta overwritten
    dg2, db2 = torch.ones_like(dg), torch.ones_like(db)
    ops.layernorm_bwd(dy, x, mean, rstd, w, None, dgamma=dg2, dbeta=db2, acc_mask=1)
    assert rel_err(dg2, wr.grad + 1) < 1e-4 and rel_err(db2, br.grad) < 1e-4


# ------------------------------------------------------------------------------------------------
# Attention
# ------------------------------------------------------------------------------------------------
def _attn_ref(qkv, B, N, H, hd, dout=None):
    D = H * hd
    q5 = qkv.float().reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4).detach().clone().requires_grad_(True)
    q, k, v = q5[0], q5[1], q5[2]
    s = q @ k.transpose(-2, -1) * hd ** -0.5
    p = torch.softmax(s, -1)
    o = (p @ v).transpose(1, 2).reshape(B * N, D)
    lse = torch.logsumexp(s, -1)
    if dout is None:
        return o, lse, None
    o.backward(dout.float())
    dqkv = q5.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    return o.detach(), lse.detach(), dqkv


@pytest.mark.parametrize("B,N,H,hd", [(3, 256, 4, 64), (5, 64, 8, 64), (6, 16, 8, 80), (7, 4, 8, 80), (2, 16, 4, 32),
                                      (2, 64, 2, 32),
                                      # short grids with several heads packed per warp (4 / 2 / 8 heads), and the shapes
                                      # where the packing does not apply (6 heads of 4 tokens, 12 tokens)
                                      (5, 4, 32, 80), (3, 8, 4, 64), (2, 2, 8, 32), (3, 4, 6, 80), (2, 12, 4, 64)])
def test_attention_fwd_bwd(B, N, H, hd):
    ops = _ops()
    D = H * hd
    qkv = randn(B * N, 3 * D, seed=30)
    dout = randn(B * N, D, seed=31)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)
    oref, lref, dref = _attn_ref(qkv, B, N, H, hd, dout)
    assert max_err_scaled(out.float(), oref) < 1.5e-2
    assert rel_err(out.float(), oref) < 8e-3
    assert float((lse - lref).abs().max()) < 2e-3
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd)
    assert rel_err(dqkv.float(), dref) < 1.5e-2
    assert max_err_scaled(dqkv.float(), dref) < 2e-2
    for part in range(3):  # q, k, v gradients separately
        sl = slice(part * D, (part + 1) * D)
        assert rel_err(dqkv[:, sl].float(), dref[:, sl]) < 2e-2, f"part {part}"


# ------------------------------------------------------------------------------------------------
# Integer index maps: bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,p", [(256, 16), (256, 32), (256, 64), (256, 128), (64, 8), (32, 16)])
def test_patch_index_maps_bit_exact(S, p):
    ops = _ops()
    B = 2
    rng = np.random.default_rng(S + p)
    imgs_np = rng.standard_normal((B, 3, S, S), dtype=np.float32)
    imgs = torch.from_numpy(imgs_np).cuda()
    pat = ops.patchify(imgs, p)
    assert np.array_equal(pat.cpu().numpy(), O.patchify_np(imgs_np, p))
    back = ops.unpatchify(pat, p)
    assert torch.equal(back, imgs)
    h = imgs.to(torch.bfloat16)
    pat16 = ops.patchify(h, p)
    assert torch.equal(pat16, O.patchify(h, p))
    assert torch.equal(ops.unpatchify(pat16, p), h)
    if p % 8 == 0:
        cols = ops.im2col(imgs, p)
        want = torch.from_numpy(O.im2col_np(imgs_np, p)).to(torch.bfloat16)
        assert torch.equal(cols.cpu(), want)


def test_patchify_integer_payload_roundtrip():
    """Indices carried as int32 payloads: the permutation itself, independent of float formatting."""
    ops = _ops()
    S, p, B = 256, 16, 1
    idx = torch.arange(B * 3 * S * S, dtype=torch.int32).reshape(B, 3, S, S).cuda()
    pat = ops.patchify(idx, p)
    want = O.patchify_np(idx.cpu().numpy(), p)
    assert np.array_equal(pat.cpu().numpy(), want)
    assert torch.equal(ops.unpatchify(pat, p), idx)


# ------------------------------------------------------------------------------------------------
# Loss, reductions, optimizer
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,p", [(2, 256, 16), (3, 64, 8), (2, 256, 128)])
def test_mse_loss_and_grad(B, S, p):
    ops = _ops()
    g = S // p
    imgs = randn(B, 3, S, S, dtype=torch.float32, seed=40)
    pred = randn(B, g * g, 3 * p * p, seed=41)
    loss, _ = ops.mse_loss(pred, imgs, p)
    pr = pred.float().clone().requires_grad_(True)
    lref = O.forward_loss(imgs, pr, p)
    assert abs(float(loss) - float(lref)) < 1e-5 * abs(float(lref))
    lref.backward()
    scale = torch.tensor([3.0], device="cuda")
    _, dpred = ops.mse_loss(pred, imgs, p, want_grad=True, grad_scale=scale)
    assert rel_err(dpred.float(), 3.0 * pr.grad) < 4e-3


def test_colsum_and_batch_sum():
    ops = _ops()
    x = randn(3000, 1544, seed=50)
    out = ops.colsum(x)
    assert rel_err(out, x.float().sum(0)) < 1e-5
    acc = torch.ones(1544, device="cuda")
    ops.colsum(x, out=acc, accumulate=True)
    assert rel_err(acc, x.float().sum(0) + 1) < 1e-5
    y = randn(5 * 64, 256, dtype=torch.float32, seed=51)
    bs = ops.batch_sum(y, 5, 64)
    assert rel_err(bs, y.reshape(5, 64, 256).sum(0)) < 1e-6


@pytest.mark.parametrize("n", [4096 * 3 + 5, 1 << 20])
def test_adamw_matches_oracle_and_torch(n):
    ops = _ops()
    p = randn(n, dtype=torch.float32, seed=60)
    g = randn(n, dtype=torch.float32, seed=61) * 0.01
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    pb = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pt], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    po, mo, vo = p.clone(), m.clone(), v.clone()
    sq = torch.zeros(1, device="cuda")
    for step in (1, 2, 3):
        pt.grad = g.clone()
        opt.step()
        po, mo, vo = O.adamw_step(po, g, mo, vo, step, 1e-3, 0.9, 0.95, 1e-8, 0.05)
        sq.zero_()
        ops.adamw_step(p, g, m, v, pb, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.05, step=step,
                       grad_sq_sum=sq)
    assert rel_err(p, pt.detach()) < 1e-6
    assert rel_err(p, po) < 1e-6
    assert torch.equal(pb, p.to(torch.bfloat16))
    assert abs(float(sq) - float((g * g).sum())) < 1e-4 * float((g * g).sum())
    # found_inf skips the update
    flag = torch.ones(1, dtype=torch.int32, device="cuda")
    before = p.clone()
    ops.adamw_step(p, g, m, v, pb, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.05, step=4, found_inf=flag)
    assert torch.equal(p, before)


# ------------------------------------------------------------------------------------------------
# fp32 mode (csrc/fp32_mode.cu, tae_b200/fp32.py): fp64 PyTorch as the reference of each op, 1e-5-level gates
# ------------------------------------------------------------------------------------------------
def _fp32():
    from tae_b200 import fp32

    return fp32


def test_fp32_split3_is_exact():
    F = _fp32()
    x = randn(300, 136, dtype=torch.float32, seed=3, scale=7.0)
    hi, mid, lo = F.split3(x)
    assert torch.equal(hi.float() + mid.float() + lo.float(), x)  # 3 x 8 mantissa bits: exact reconstruction
    assert torch.equal(hi, x.to(torch.bfloat16))


@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(300, 136, 264, False, False), (520, 128, 1024, False, True),
                                             (128, 72, 4104, True, True)])
def test_fp32_split_gemm(M, N, K, a_mn, b_mn):
    F = _fp32()
    A = randn(M, K, dtype=torch.float32, seed=1)
    Bm = randn(N, K, dtype=torch.float32, seed=2)
    ref = (A.double() @ Bm.double().t())
    As = A.t().contiguous() if a_mn else A
    Bs = Bm.t().contiguous() if b_mn else Bm
    out = F.gemm_f32(F.split3(As), F.split3(Bs), a_mn=a_mn, b_mn=b_mn)
    # tensor-core fp32 accumulation truncates: error grows slowly with K (4e-6 at K=4104)
    assert out.dtype == torch.float32 and rel_err(out, ref) < 1e-5
    # accumulate form used by the weight gradients
    out2 = F.gemm_f32(F.split3(As), F.split3(Bs), a_mn=a_mn, b_mn=b_mn, out=out.clone(), beta=1)
    assert rel_err(out2, 2 * ref) < 1e-5
    # plain bf16 GEMM on the same data is ~3 orders of magnitude coarser: the split really buys fp32 accuracy
    coarse = _ops().gemm(As.to(torch.bfloat16), Bs.to(torch.bfloat16), a_mn=a_mn, b_mn=b_mn,
                         epilogue=3)  # TAE_EPI_F32_ACC
    assert rel_err(coarse, ref) > 1e-4


@pytest.mark.parametrize("rows,D", [(77, 128), (512, 1024), (33, 2560)])
def test_fp32_layernorm_fwd_bwd(rows, D):
    F = _fp32()
    x = randn(rows, D, dtype=torch.float32, seed=4, scale=2.0)
    w = randn(D, dtype=torch.float32, seed=5) * 0.5 + 1.0
    b = randn(D, dtype=torch.float32, seed=6)
    dy = randn(rows, D, dtype=torch.float32, seed=7)
    dres = randn(rows, D, dtype=torch.float32, seed=8)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-6)
    yr.backward(dy.double())
    y, mean, rstd = F.layernorm_fwd(x, w, b, 1e-6)
    assert rel_err(y, yr.detach()) < 1e-6
    dx, dg, db = F.layernorm_bwd(dy, x, mean, rstd, w, dres)
    assert rel_err(dx, xr.grad + dres.double()) < 1e-5
    assert rel_err(dg, wr.grad) < 1e-5 and rel_err(db, br.grad) < 1e-5
    dg2, db2 = dg.clone(), db.clone()
    F.layernorm_bwd(dy, x, mean, rstd, w, None, dg2, db2, acc_mask=3)
    assert rel_err(dg2, 2 * wr.grad) < 1e-5 and rel_err(db2, 2 * br.grad) < 1e-5


@pytest.mark.parametrize("B,N,H,hd", [(2, 256, 2, 64), (3, 64, 4, 32), (2, 16, 2, 80), (5, 4, 2, 80)])
def test_fp32_attention_fwd_bwd(B, N, H, hd):
    F = _fp32()
    D = H * hd
    qkv = randn(B * N, 3 * D, dtype=torch.float32, seed=9, scale=0.7)
    dout = randn(B * N, D, dtype=torch.float32, seed=10)
    qr = qkv.double().requires_grad_(True)
    q, k, v = qr.view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, D)
    o.backward(dout.double())
    out, lse = F.attention_fwd(qkv, B, N, H, hd)
    assert rel_err(out, o.detach()) < 1e-5
    lse_ref = torch.logsumexp((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    assert rel_err(lse, lse_ref.detach()) < 1e-5
    dqkv = F.attention_bwd(qkv, out, dout, lse, B, N, H, hd)
    assert rel_err(dqkv, qr.grad) < 2e-5


def test_fp32_bias_gelu_loss_im2col():
    F = _fp32()
    M, N = 130, 264
    acc = randn(M, N, dtype=torch.float32, seed=11)
    bias = randn(N, dtype=torch.float32, seed=12)
    pos = randn(26, N, dtype=torch.float32, seed=13)
    y = acc.clone()
    act = F.bias_act(y, bias, pos, 26, want_act=True)
    yr = acc.double() + bias.double() + pos.double().repeat(5, 1)
    assert rel_err(y, yr) < 1e-6 and rel_err(act, torch.nn.functional.gelu(yr)) < 1e-6
    hr = yr.clone().requires_grad_(True)
    da = randn(M, N, dtype=torch.float32, seed=14)
    torch.nn.functional.gelu(hr).backward(da.double())
    assert rel_err(F.gelu_bwd(y, da), hr.grad) < 1e-5
    # loss + gradient and im2col against the oracle's restatement
    imgs = randn(3, 3, 64, 64, dtype=torch.float32, seed=15)
    pred = randn(3, 16, 768, dtype=torch.float32, seed=16)
    pr = pred.double().requires_grad_(True)
    lr = ((pr - O.patchify(imgs.double(), 16)) ** 2).mean()
    lr.backward()
    gs = torch.full((1,), 3.0, device="cuda")
    loss, dpred = F.mse_loss(pred, imgs, 16, want_grad=True, grad_scale=gs)
    assert abs(float(loss) - float(lr)) < 1e-6 * float(lr) and rel_err(dpred, 3.0 * pr.grad) < 1e-6
    cols = F.im2col(imgs, 16)
    assert torch.equal(cols, O.im2col(imgs, 16).reshape(cols.shape))


# ------------------------------------------------------------------------------------------------
# ROWDOT epilogue (delta = rowsum(dO * O) emitted by the proj dgrad) and attention backward fed with it
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,tokens", [(512, 128, 264, 256), (1024, 1024, 1024, 256), (300, 64, 4104, 100)])
def test_gemm_rowdot_epilogue(M, N, K, tokens):
    ops = _ops()
    A, W = randn(M, K, seed=21, scale=0.5), randn(K, N, seed=22, scale=0.5)   # dgrad layout: B is [K, N] (b_mn)
    aux = randn(M, N, seed=23)
    out, rd = ops.gemm(A, W, b_mn=True, epilogue=5, aux=aux, rowdot_tokens=tokens)
    ref = (A.float() @ W.float()).to(torch.bfloat16)
    assert max_err_scaled(out.float(), ref.float()) < 1e-2
    # the row-dot is taken over the ROUNDED outputs the kernel wrote
    want = (out.float() * aux.float()).view(M // tokens, tokens, N // 64, 64).sum(-1).permute(0, 2, 1).contiguous()
    assert rd.shape == (M // tokens, N // 64, tokens) and rel_err(rd, want) < 1e-5


def test_attention_bwd_with_precomputed_delta():
    ops = _ops()
    B, N, H, hd = 5, 256, 3, 64
    D = H * hd
    qkv = randn(B * N, 3 * D, seed=24, scale=0.7)
    dout = randn(B * N, D, seed=25)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)
    ref = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd)
    delta = (dout.float() * out.float()).view(B, N, H, hd).sum(-1).permute(0, 2, 1).contiguous()
    got = ops.attention_bwd(qkv, None, dout, lse, B, N, H, hd, delta=delta)
    assert max_err_scaled(got.float(), ref.float()) < 2e-3


@pytest.mark.parametrize("B,H", [(3, 2), (37, 32), (1000, 1)])
def test_attention_n64_kernels_with_and_without_delta(B, H):
    """The 64-token grid (patch32): persistent mma.sync kernels, several items per CTA (1184 and 1000 items over 444 / 296
    CTAs), backward both with O staged (delta computed in the kernel, 2 CTAs per SM) and with the precomputed delta of the
    row-dot GEMM epilogue (3 CTAs per SM); dS^T goes through shared memory between the two phases."""
    ops = _ops()
    N, hd = 64, 64
    D = H * hd
    qkv = randn(B * N, 3 * D, seed=50 + H, scale=0.8)
    dout = randn(B * N, D, seed=51 + H)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)
    oref, lref, dref = _attn_ref(qkv, B, N, H, hd, dout)
    assert rel_err(out.float(), oref) < 8e-3 and max_err_scaled(out.float(), oref) < 1.5e-2
    assert float((lse - lref).abs().max()) < 2e-3
    delta = (dout.float() * out.float()).view(B, N, H, hd).sum(-1).permute(0, 2, 1).contiguous()
    got = []
    for kw in ({}, {"delta": delta}):
        dqkv = ops.attention_bwd(qkv, None if kw else out, dout, lse, B, N, H, hd, **kw)
        assert rel_err(dqkv.float(), dref) < 1.5e-2 and max_err_scaled(dqkv.float(), dref) < 2e-2
        for part in range(3):  # q, k, v gradients separately
            sl = slice(part * D, (part + 1) * D)
            assert rel_err(dqkv[:, sl].float(), dref[:, sl]) < 2e-2, f"part {part}"
        per_img = ((dqkv.float() - dref).view(B, -1).norm(dim=1) / dref.view(B, -1).norm(dim=1)).max()
        assert float(per_img) < 2e-2
        got.append(dqkv)
    assert max_err_scaled(got[1].float(), got[0].float()) < 2e-3
    assert torch.equal(ops.attention_bwd(qkv, None, dout, lse, B, N, H, hd, delta=delta), got[1])  # deterministic


@pytest.mark.parametrize("B,H", [(41, 16), (300, 1)])
def test_attention_persistent_kernels_many_items_per_cta(B, H):
    """The tcgen05 attention kernels are persistent: with more (image, head) items than SMs every CTA walks several
    items, which exercises the cross-item operand prefetch, the double-buffered lse/delta and operand homes and the
    barrier phase bookkeeping (item counts per CTA differ: 656 and 300 items over 148 CTAs)."""
    ops = _ops()
    N, hd = 256, 64
    D = H * hd
    qkv = randn(B * N, 3 * D, seed=40 + H, scale=0.8)
    dout = randn(B * N, D, seed=41 + H)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)
    oref, lref, dref = _attn_ref(qkv, B, N, H, hd, dout)
    assert rel_err(out.float(), oref) < 8e-3 and max_err_scaled(out.float(), oref) < 1.5e-2
    assert float((lse - lref).abs().max()) < 2e-3
    # per-image error: a wrong item mapping or a stale operand buffer would show up in single images
    per_img = ((out.float() - oref).view(B, -1).norm(dim=1) / oref.view(B, -1).norm(dim=1)).max()
    assert float(per_img) < 1e-2
    for kw in ({}, {"delta": (dout.float() * out.float()).view(B, N, H, hd).sum(-1).permute(0, 2, 1).contiguous()}):
        dqkv = ops.attention_bwd(qkv, None if kw else out, dout, lse, B, N, H, hd, **kw)
        assert rel_err(dqkv.float(), dref) < 1.5e-2 and max_err_scaled(dqkv.float(), dref) < 2e-2
        per_img = ((dqkv.float() - dref).view(B, -1).norm(dim=1) / dref.view(B, -1).norm(dim=1)).max()
        assert float(per_img) < 2e-2
    # determinism: the kernels use no atomics
    out2, lse2 = ops.attention_fwd(qkv, B, N, H, hd)
    assert torch.equal(out2, out) and torch.equal(lse2, lse)
    assert torch.equal(ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd), ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd))


def test_dynamic_scheduling_matches_static():
    """tae_set_dynamic_scheduling: the persistent kernels draw their work (GEMM tiles, LayerNorm-backward rows) from a
    global counter instead of static round-robin lists (tae_b200.ddp turns this on for world_size > 1).  Which CTA
    computes a tile must not change any value: every per-element result is bit-identical (only sums
    whose order follows the work assignment — split-K atomics, LayerNorm parameter-gradient partials — may differ in the
    last bits)."""
    ops = _ops()
    from tae_b200._lib import EPI_BF16, EPI_BF16_GELU, EPI_F32_ACC

    A, W = randn(8192, 1024, seed=50), randn(2048, 1024, seed=51, scale=0.05)
    bias = randn(2048, dtype=torch.float32, seed=52)
    dY, X = randn(8192, 640, seed=53, scale=0.1), randn(8192, 384, seed=54)
    x = randn(5000, 1024, dtype=torch.float32, seed=55) * 2 + 0.5
    w = randn(1024, dtype=torch.float32, seed=56) * 0.2 + 1
    dy, dres = randn(5000, 1024, seed=57), randn(5000, 1024, dtype=torch.float32, seed=58)
    _, mean, rstd = ops.layernorm_fwd(x, w, torch.zeros_like(w), 1e-6)
    # the wide-row LayerNorm backward (row groups, D = 2048 / 2560) draws its rows from the same counter slots
    wide = []
    for rows_w, Dw in ((3001, 2048), (1500, 2560)):
        xw = randn(rows_w, Dw, dtype=torch.float32, seed=59) * 2 + 0.5
        ww = randn(Dw, dtype=torch.float32, seed=60) * 0.2 + 1
        _, mw, rw = ops.layernorm_fwd(xw, ww, torch.zeros_like(ww), 1e-6)
        wide.append((randn(rows_w, Dw, seed=61), xw, mw, rw, ww, randn(rows_w, Dw, dtype=torch.float32, seed=62)))

    def run():
        plain = ops.gemm(A, W, epilogue=EPI_BF16, bias=bias)
        gp, g = ops.gemm(A, W, epilogue=EPI_BF16_GELU, bias=bias)
        dw = ops.gemm(dY, X, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC, splits=1)
        dw_split = ops.gemm(dY, X, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC, splits=0)
        ln = ops.layernorm_bwd(dy, x, mean, rstd, w, dres)
        lw = [ops.layernorm_bwd(*a) for a in wide]
        torch.cuda.synchronize()
        # exact: per-element results; approx: sums whose order follows the row -> CTA assignment or the split-K atomics
        return ([plain, gp, g, dw, ln[0], ln[1]] + [t for l in lw for t in l[:2]],
                [dw_split, *[t for t in ln[2:] if t is not None]] + [t for l in lw for t in l[2:] if t is not None])

    prev = ops.set_dynamic_scheduling(False)
    try:
        static, static_sums = run()
        assert ops.set_dynamic_scheduling(True) == 0
        for _ in range(3):  # several launches: the counter slots are re-armed by the kernels themselves
            dynamic, dynamic_sums = run()
            for s, d in zip(static, dynamic):
                assert torch.equal(s, d)
            for s, d in zip(static_sums, dynamic_sums):
                assert rel_err(d, s) < 1e-5
        assert ops.set_dynamic_scheduling(False) == 1
    finally:
        if prev >= 0:
            ops.set_dynamic_scheduling(bool(prev))
        else:
            ops.set_dynamic_scheduling(False)
    ref = (A.float() @ W.float().t() + bias.to(torch.bfloat16).float())
    assert max_err_scaled(static[0].float(), ref) < 1e-2


@pytest.mark.parametrize("B,H,kscale_hi", [(3, 2, 1.0), (2, 3, 12.0), (5, 1, 0.05)])
def test_attention_fwd_two_key_halves(B, H, kscale_hi):
    """Forward attention at N = 256 when the two 128-key halves of a row have very different score ranges: the kernel
    carries the first half's row maximum into the second half and rescales the accumulator only when the second half's
    maximum exceeds it by 2^32.
    kscale_hi = 12: most rows take the rescale path; 0.05: the second half is negligible; 1: the common case."""
    ops = _ops()
    N, hd = 256, 64
    D = H * hd
    qkv = randn(B * N, 3 * D, seed=90)
    k = qkv.view(B, N, 3, H, hd)[:, 128:, 1]
    k.mul_(kscale_hi)  # keys 128..255 of every image and head
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)
    oref, lref, _ = _attn_ref(qkv, B, N, H, hd)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert max_err_scaled(out.float(), oref) < 1.5e-2 and rel_err(out.float(), oref) < 8e-3
    assert float((lse - lref).abs().max()) < 2e-3 * max(1.0, kscale_hi)
