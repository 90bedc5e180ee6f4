"""GPU parity of the downstream ViTs (reference tae.py:274-429) against fixtures generated from the unmodified reference
(tests/golden/make_golden_vit.py): outputs, loss and every stored gradient; bf16 path within 2e-2 (gradients too), fp32 mode within 1e-4
(2e-4 on gradients)."""
import json
import os
from functools import partial

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from conftest import GOLDEN_DIR, report  # noqa: E402

CASES = ["vitrec_n16_hd32_c37", "vitrec_n256_hd64_c16", "vitseg_n16_p8_c5"]


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def vit_meta():
    with open(os.path.join(GOLDEN_DIR, "vit_golden_meta.json")) as f:
        return json.load(f)


def loss_of(cls, out, labels):
    if cls == "VITForRecognition":
        return F.cross_entropy(out.float(), labels)
    return F.cross_entropy(out["out"].float(), labels) + 0.5 * F.cross_entropy(out["aux"].float(), labels)


@pytest.mark.parametrize("mode,tol,gtol", [("bf16", 2e-2, 2e-2), ("fp32", 1e-4, 2e-4)])
@pytest.mark.parametrize("case", CASES)
def test_vit_forward_backward_matches_reference(case, mode, tol, gtol, vit_meta):
    from tae_b200 import tae as T

    rec = vit_meta[case]
    t = torch.load(os.path.join(GOLDEN_DIR, f"{case}.pt"), map_location="cpu")
    torch.manual_seed(0)
    model = getattr(T, rec["class"])(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **rec["kwargs"]).cuda().train()
    model.set_precision(mode)
    out = model(t["input"].cuda())
    loss = loss_of(rec["class"], out, t["labels"].cuda())
    loss.backward()
    torch.cuda.synchronize()
    g = rec[mode]
    outs = {"out": out} if rec["class"] == "VITForRecognition" else out
    for k, v in outs.items():
        assert str(v.dtype) == g["out_dtype"][k], (k, v.dtype)           # the reference's output dtypes
        assert tuple(v.shape) == tuple(t[f"{mode}.{k}"].shape)
        assert rel(v.float().cpu(), t[f"{mode}.{k}"]) < tol, k
    assert abs(float(loss) - g["loss"]) < tol * abs(g["loss"])
    vs_norm, vs_ref = [], []
    for n, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, n
        gn = float(p.grad.norm())
        vs_norm.append((n, abs(gn - g["grad_norm"][n]) / (g["grad_norm"][n] + 1e-7)))
        key = f"{mode}.grad.{n}"
        if key in t and float(t[key].norm()) > 1e-6:
            vs_ref.append((n, rel(p.grad.cpu(), t[key])))
    worst = [report(f"{case} {mode} grad norm vs reference", vs_norm), report(f"{case} {mode} grad vs reference", vs_ref)]
    assert all(w[1] < gtol for w in worst), worst


def test_vit_recognition_features_and_headless(vit_meta):
    from tae_b200 import tae as T

    rec = vit_meta["vitrec_n16_hd32_c37"]
    kw = dict(rec["kwargs"], num_classes=None)
    torch.manual_seed(0)
    m = T.VITForRecognition(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw).cuda().eval()
    z = torch.randn(4, 16, 64, device="cuda")
    with torch.no_grad():
        f = m.forward_features(z)
        y = m(z)
    assert f.dtype == torch.float32 and f.shape == (4, 16, 128)
    assert y.shape == (4, 128) and rel(y, f.mean(dim=1)) < 1e-6


def test_unpatchify_c_is_the_reference_permutation():
    from tae_b200 import ops

    B, g, p, C = 2, 4, 8, 5
    x = torch.randn(B, g * g, p * p * C, device="cuda")
    ref = torch.einsum("nhwpqc->nchpwq", x.reshape(B, g, g, p, p, C)).reshape(B, C, g * p, g * p)
    y = ops.unpatchify_c(x, p, C)
    assert torch.equal(y, ref)                                   # integer index map: bit-exact
    assert torch.equal(ops.patchify_c(y, p), x)                  # and its inverse
    xb = x.to(torch.bfloat16)
    assert torch.equal(ops.unpatchify_c(xb, p, C), ref.to(torch.bfloat16))
