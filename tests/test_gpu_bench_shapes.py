"""Per-kernel parity AT THE BENCHMARKED SHAPES (BASELINE configs 2 and 4): M = 65 536 token rows, N = 49 152
(`decoder_pred` of patch128), K = 49 152 with split-K (PatchEmbed weight gradient of patch128), attention and LayerNorm
over the full B = 256 batch.  These are the sizes where 32-bit offset arithmetic would break (M*N = 2^28 elements,
byte offsets past 2^31) and that the small kernel tests never reach.  Checker = the oracle's restatement of each op in
fp32 on the same device; each case takes milliseconds."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tae_oracle as O  # noqa: E402  (checker only)


def _ops():
    from tae_b200 import ops

    return ops


def rel_err(a, b):
    # chunked so that no fp64 copy of a 2^28-element tensor is made
    num = den = 0.0
    a, b = a.reshape(-1), b.reshape(-1)
    step = 1 << 26
    for i in range(0, a.numel(), step):
        x, y = a[i:i + step].double(), b[i:i + step].double()
        num += float((x - y).pow(2).sum())
        den += float(y.pow(2).sum())
    return math.sqrt(num / max(den, 1e-60))


def randn(*shape, dtype=torch.bfloat16, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


M16 = 65536  # B=256 x 256 tokens


@pytest.mark.parametrize("N,K,epi", [(3072, 1024, "bf16"), (4096, 1024, "gelu"), (1024, 4096, "resid"), (1024, 1024, "resid"),
                                     (256, 1024, "bf16"), (768, 1024, "bf16")])
def test_gemm_forward_epilogues_at_patch16_bench_shape(N, K, epi):
    ops = _ops()
    from tae_b200._lib import EPI_BF16, EPI_BF16_GELU, EPI_F32_RESID

    A, W = randn(M16, K, seed=1), randn(N, K, seed=2, scale=0.05)
    bias = randn(N, dtype=torch.float32, seed=3)
    acc = A.float() @ W.float().t() + bias.to(torch.bfloat16).float()
    if epi == "bf16":
        out = ops.gemm(A, W, epilogue=EPI_BF16, bias=bias)
        assert rel_err(out.float(), acc) < 5e-3
    elif epi == "gelu":
        gp, a = ops.gemm(A, W, epilogue=EPI_BF16_GELU, bias=bias)
        hf = acc.to(torch.bfloat16).float()
        assert rel_err(a.float(), O.gelu(hf)) < 4e-3
        gpref = 0.5 * (1 + torch.erf(hf / math.sqrt(2))) + hf * torch.exp(-0.5 * hf * hf) / math.sqrt(2 * math.pi)
        assert rel_err(gp.float(), gpref) < 4e-3
    else:
        resid = randn(M16, N, dtype=torch.float32, seed=4)
        out = ops.gemm(A, W, epilogue=EPI_F32_RESID, bias=bias, resid=resid)
        y = acc.to(torch.bfloat16).float()
        assert rel_err(out - resid, y) < 5e-3
        # the last rows are where a 32-bit element offset (M*N*4 bytes > 2^31 for N >= 8192 only, but M*N*4 = 2^28 here)
        # or a wrong tile walk would show first
        assert rel_err((out - resid)[-300:], y[-300:]) < 5e-3


@pytest.mark.parametrize("Nout,Kin", [(1024, 4096), (4096, 1024), (3072, 1024)])
def test_gemm_dgrad_wgrad_at_patch16_bench_shape(Nout, Kin):
    ops = _ops()
    from tae_b200._lib import EPI_BF16, EPI_F32_ACC

    dY, X, W = randn(M16, Nout, seed=5, scale=0.1), randn(M16, Kin, seed=6), randn(Nout, Kin, seed=7, scale=0.05)
    dX = ops.gemm(dY, W, b_mn=True, epilogue=EPI_BF16)                       # [M, Kin] = dY W
    assert rel_err(dX.float(), dY.float() @ W.float()) < 5e-3
    dW = ops.gemm(dY, X, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)          # [Nout, Kin] = dY^T X, K = 65 536
    assert rel_err(dW, dY.float().t() @ X.float()) < 1e-3


def test_gemm_rowdot_and_dgelu_at_patch16_bench_shape():
    ops = _ops()
    from tae_b200._lib import EPI_BF16_DGELU, EPI_BF16_ROWDOT

    D = 1024
    dY, W, att = randn(M16, D, seed=8, scale=0.1), randn(D, D, seed=9, scale=0.05), randn(M16, D, seed=10)
    datt, delta = ops.gemm(dY, W, b_mn=True, epilogue=EPI_BF16_ROWDOT, aux=att, rowdot_tokens=256)
    ref = dY.float() @ W.float()
    assert rel_err(datt.float(), ref) < 5e-3
    want = (datt.float() * att.float()).view(256, 256, 16, 64).sum(-1).permute(0, 2, 1)
    assert rel_err(delta, want) < 1e-5
    del ref, want, att
    W2, gp = randn(D, 4 * D, seed=11, scale=0.05), randn(M16, 4 * D, seed=12, scale=0.5)
    part = torch.empty(M16 // 32, 4 * D, device="cuda")
    dh = ops.gemm(dY, W2, b_mn=True, epilogue=EPI_BF16_DGELU, aux=gp, colsum_partials=part)
    ref = (dY.float() @ W2.float()).to(torch.bfloat16).float() * gp.float()
    assert rel_err(dh.float(), ref) < 5e-3
    assert rel_err(ops.colsum_f32(part), dh.float().sum(0)) < 1e-4


def test_gemm_patch128_shapes_wide_n_and_split_k():
    """config 4: M = 1024 rows; decoder_pred N = 49 152; PatchEmbed K = 49 152 (forward) and its weight gradient."""
    ops = _ops()
    from tae_b200._lib import EPI_BF16, EPI_F32_ACC, EPI_F32_RESID

    M, D, KP = 1024, 2560, 49152
    x, Wp = randn(M, D, seed=13), randn(KP, D, seed=14, scale=0.02)
    bias = randn(KP, dtype=torch.float32, seed=15)
    pred = ops.gemm(x, Wp, epilogue=EPI_BF16, bias=bias)                      # [1024, 49152]
    assert rel_err(pred.float(), x.float() @ Wp.float().t() + bias.to(torch.bfloat16).float()) < 5e-3
    cols, We = randn(M, KP, seed=16), randn(D, KP, seed=17, scale=0.01)
    pos = randn(4, D, dtype=torch.float32, seed=18)
    emb = ops.gemm(cols, We, epilogue=EPI_F32_RESID, resid=pos, resid_rows=4)  # K = 49152
    ref = (cols.float() @ We.float().t()).to(torch.bfloat16).float() + pos.repeat(M // 4, 1)
    assert rel_err(emb, ref) < 5e-3
    dY = randn(M, D, seed=19, scale=0.1)
    dW = ops.gemm(dY, cols, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)        # [2560, 49152], split-K over 1024 rows
    assert rel_err(dW, dY.float().t() @ cols.float()) < 1e-3
    dWp = ops.gemm(pred, x, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)        # [49152, 2560]
    assert rel_err(dWp, pred.float().t() @ x.float()) < 1e-3
    dx = ops.gemm(pred, Wp, b_mn=True, epilogue=EPI_BF16)                     # dgrad with K = 49152
    assert rel_err(dx.float(), pred.float() @ Wp.float()) < 5e-3


def test_layernorm_at_bench_shapes():
    ops = _ops()
    for rows, D in ((M16, 1024), (1024, 2560), (16384, 2048)):
        x = randn(rows, D, dtype=torch.float32, seed=20) * 2 + 0.5
        w = randn(D, dtype=torch.float32, seed=21) * 0.2 + 1
        b = randn(D, dtype=torch.float32, seed=22) * 0.1
        y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
        xr = x.clone().requires_grad_(True)
        wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = O.layer_norm(xr, wr, br, 1e-6)
        assert rel_err(y.float(), yr.detach()) < 4e-3
        dy = randn(rows, D, seed=23)
        dres = randn(rows, D, dtype=torch.float32, seed=24)
        yr.backward(dy.float())
        dx, dx_b, dg, db, cs = ops.layernorm_bwd(dy, x, mean, rstd, w, dres)
        assert rel_err(dx - dres, xr.grad) < 1e-4
        assert rel_err(dg, wr.grad) < 2e-4 and rel_err(db, br.grad) < 2e-4
        assert torch.equal(dx_b, dx.to(torch.bfloat16))
        assert rel_err(cs, dx_b.float().sum(0)) < 2e-4
        del xr, yr, dx, dx_b


@pytest.mark.parametrize("B,N,H,hd", [(256, 256, 16, 64), (256, 64, 32, 64), (256, 16, 32, 80), (256, 4, 32, 80)])
def test_attention_at_bench_shapes(B, N, H, hd):
    ops = _ops()
    D = H * hd
    qkv = randn(B * N, 3 * D, seed=30)
    dout = randn(B * N, D, seed=31)
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd)
    # reference in chunks of 32 images (the fp32 score tensor of the whole batch would be 4 GB at N = 256)
    errs_o, errs_d, nrm_o, nrm_d = 0.0, 0.0, 0.0, 0.0
    for b0 in range(0, B, 32):
        rows = slice(b0 * N, (b0 + 32) * N)
        q5 = qkv[rows].float().reshape(32, N, 3, H, hd).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
        s = q5[0] @ q5[1].transpose(-2, -1) * hd ** -0.5
        o = (torch.softmax(s, -1) @ q5[2]).transpose(1, 2).reshape(32 * N, D)
        o.backward(dout[rows].float())
        dref = q5.grad.permute(1, 3, 0, 2, 4).reshape(32 * N, 3 * D)
        errs_o += float((out[rows].float() - o.detach()).double().pow(2).sum())
        nrm_o += float(o.detach().double().pow(2).sum())
        errs_d += float((dqkv[rows].float() - dref).double().pow(2).sum())
        nrm_d += float(dref.double().pow(2).sum())
        assert float((lse.reshape(B, H, N)[b0:b0 + 32] - torch.logsumexp(s, -1).detach()).abs().max()) < 2e-3
    assert math.sqrt(errs_o / nrm_o) < 8e-3
    assert math.sqrt(errs_d / nrm_d) < 1.5e-2


def test_loss_and_im2col_at_bench_shape():
    ops = _ops()
    B = 256
    imgs = randn(B, 3, 256, 256, dtype=torch.float32, seed=40)
    for p in (16, 128):
        pred = randn(B, (256 // p) ** 2, 3 * p * p, seed=41)
        loss, dpred = ops.mse_loss(pred, imgs, p, want_grad=True)
        tgt = O.patchify(imgs, p)
        diff = pred.float() - tgt
        assert abs(float(loss) - float(diff.pow(2).mean())) < 1e-5 * float(loss)
        assert rel_err(dpred.float(), 2 * diff / diff.numel()) < 4e-3
        cols = ops.im2col(imgs, p)
        g = 256 // p
        want = imgs.reshape(B, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, 3 * p * p).to(torch.bfloat16)
        assert torch.equal(cols, want)  # integer index map: bit-exact
