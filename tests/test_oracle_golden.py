"""CPU tests: the oracle (oracle/tae_oracle.py) against the golden fixtures generated from the unmodified reference
(tests/golden/make_golden.py), plus the seed-pinned scalars of the real patch16 model."""
from functools import partial

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, oracle_cfg
from oracle import tae_oracle as O


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model_state(kw):
    """Weights come from tae_b200's seeded init, which must reproduce the reference's (checksums in the fixture)."""
    from tae_b200 import tae as T

    torch.manual_seed(0)
    m = T.TAE(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_seeded_init_matches_reference_checksums(case, golden_meta):
    rec = golden_meta[case]
    m, sd = _model_state(rec["kwargs"])
    assert list(sd.keys()) == rec["param_names"]
    for n, (s, a) in rec["init_checksums"].items():
        assert abs(float(sd[n].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), n
        assert abs(float(sd[n].double().abs().sum()) - a) <= 1e-9 * max(1.0, a), n


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_fp32_matches_reference(case, golden_meta, golden_tensors):
    rec, t = golden_meta[case], golden_tensors(case)
    _, sd = _model_state(rec["kwargs"])
    cfg = oracle_cfg(rec["kwargs"])
    acts = {}
    loss, pred, latent = O.forward(sd, t["input"], cfg, "fp32", acts)
    g = rec["fp32"]
    assert abs(float(loss) - g["loss"]) < 1e-5 * g["loss"]
    assert rel(pred, t["fp32.pred"]) < 1e-4 and rel(latent, t["fp32.latent"]) < 1e-4
    for k, v in g["block_out_norms"].items():
        assert abs(float(acts[k].norm()) - v) < 1e-4 * v, k
    loss2, _, _, grads = O.forward_backward(sd, t["input"], cfg, "fp32")
    for k, n in enumerate(rec["param_names"]):
        gn = g["grad_norm"][n]
        assert abs(float(grads[n].norm()) - gn) < 1e-4 * gn + 1e-9, n
        probe = torch.randn(grads[n].shape, generator=torch.Generator().manual_seed(rec["seeds"]["probe"] + k))
        dot = float((grads[n].double() * probe.double()).sum())
        scale = gn * float(probe.norm()) + 1e-12
        assert abs(dot - g["grad_probe"][n]) < 2e-4 * scale, n
        key = f"fp32.grad.{n}"
        if key in t and float(t[key].norm()) > 1e-9:
            assert rel(grads[n], t[key]) < 1e-4, n
    assert abs(float(O.grad_norm(grads.values())) - g["global_grad_norm"]) < 1e-4 * g["global_grad_norm"]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_bf16_matches_reference_autocast(case, golden_meta, golden_tensors):
    rec, t = golden_meta[case], golden_tensors(case)
    _, sd = _model_state(rec["kwargs"])
    cfg = oracle_cfg(rec["kwargs"])
    loss, pred, latent, grads = O.forward_backward(sd, t["input"], cfg, "bf16")
    g = rec["bf16"]
    assert pred.dtype == torch.bfloat16 and latent.dtype == torch.bfloat16 and loss.dtype == torch.float32
    assert g["pred_dtype"] == "torch.bfloat16" and g["latent_dtype"] == "torch.bfloat16" and g["loss_dtype"] == "torch.float32"
    assert abs(float(loss) - g["loss"]) < 2e-3 * g["loss"]
    assert rel(pred.float(), t["bf16.pred"]) < 2e-2 and rel(latent.float(), t["bf16.latent"]) < 2e-2
    for n in rec["param_names"]:
        gn = g["grad_norm"][n]
        assert abs(float(grads[n].float().norm()) - gn) < 2e-2 * gn + 1e-9, n


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_adamw_matches_reference_step(case, golden_meta, golden_tensors):
    rec, t = golden_meta[case], golden_tensors(case)
    _, sd = _model_state(rec["kwargs"])
    cfg = oracle_cfg(rec["kwargs"])
    _, _, _, grads = O.forward_backward(sd, t["input"], cfg, "fp32")
    no_decay, decay = O.add_weight_decay_names([(k, tuple(v.shape)) for k, v in sd.items()], 0.05)
    assert no_decay == rec["no_decay_names"]
    a = rec["adamw"]
    for n in rec["param_names"]:
        wd = 0.0 if n in no_decay else a["weight_decay"]
        p1, _, _ = O.adamw_step(sd[n], grads[n], torch.zeros_like(sd[n]), torch.zeros_like(sd[n]), 1, a["lr"],
                                a["betas"][0], a["betas"][1], 1e-8, wd)
        upd = float((p1 - sd[n]).norm())
        assert abs(upd - a["update_norm"][n]) < 2e-3 * a["update_norm"][n] + 1e-9, n
        assert abs(float(p1.norm()) - a["param_norm_after"][n]) < 1e-5 * a["param_norm_after"][n] + 1e-9, n


@pytest.mark.parametrize("S,p", [(256, 16), (256, 32), (256, 64), (256, 128), (64, 8)])
def test_index_maps_integer_exact(S, p):
    """patchify / unpatchify / im2col as pure permutations on integer payloads, against the einsum of tae.py:196-222
    restated with torch.einsum and against conv2d for the im2col K-order."""
    n = 2
    idx = np.arange(n * 3 * S * S, dtype=np.int64).reshape(n, 3, S, S)
    g = S // p
    t = torch.from_numpy(idx)
    ref = torch.einsum("nchpwq->nhwpqc", t.reshape(n, 3, g, p, g, p)).reshape(n, g * g, p * p * 3)
    assert np.array_equal(O.patchify_np(idx, p), ref.numpy())
    assert torch.equal(O.patchify(t, p), ref)
    assert np.array_equal(O.unpatchify_np(O.patchify_np(idx, p), p), idx)
    assert torch.equal(O.unpatchify(ref, p), t)
    back = torch.einsum("nhwpqc->nchpwq", ref.reshape(n, g, g, p, p, 3)).reshape(n, 3, S, S)
    assert torch.equal(back, t)
    # im2col K-order (c,i,j): conv2d with a one-hot kernel picks cols[:, k]
    x = torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(5))
    cols = O.im2col(x, p)
    for k in (0, p * p - 1, p * p + 3, 3 * p * p - 1):
        w = torch.zeros(1, 3 * p * p)
        w[0, k] = 1.0
        y = torch.nn.functional.conv2d(x, w.reshape(1, 3, p, p), stride=p).flatten(2).transpose(1, 2)
        assert torch.equal(y[0, :, 0], cols[0, :, k])
    assert np.array_equal(O.im2col_np(idx, p).reshape(n, g * g, -1), O.im2col(t, p).numpy())


def test_pinned_scalars_real_patch16_model(golden_meta):
    """BASELINE.md §4 / SURVEY.md §8c: torch.manual_seed(0) -> tae_patch16_vocab16_px256(); seeded input; fp32 CPU."""
    from tae_b200 import tae as T

    g = golden_meta["tae_patch16_vocab16_px256_b2"]
    torch.manual_seed(0)
    m = T.tae_patch16_vocab16_px256()
    sd = m.state_dict()
    assert list(sd.keys()) == g["state_dict_keys"]
    for n, (s, a) in g["init_checksums"].items():
        assert abs(float(sd[n].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), n
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(1234))
    cfg = O.zoo_config("tae_patch16_vocab16_px256")
    loss, pred, latent, grads = O.forward_backward({k: v for k, v in sd.items()}, x, cfg, "fp32")
    assert abs(float(loss) - g["loss"]) < 1e-4 * g["loss"]
    assert abs(float(loss) - 2.184759855) < 1e-4 * 2.184759855          # the survey-time pin
    assert abs(float(pred.abs().mean()) - g["pred_abs_mean"]) < 1e-4 * g["pred_abs_mean"]
    assert abs(float(latent.abs().mean()) - g["latent_abs_mean"]) < 1e-4 * g["latent_abs_mean"]
    gn = float(O.grad_norm(grads.values()))
    assert abs(gn - g["global_grad_norm"]) < 1e-3 * g["global_grad_norm"]
