"""Model-level parity on the GPU: the CUDA path against (a) the golden fixtures generated from the unmodified
reference (tests/golden/make_golden.py) and (b) the oracle on the same weights and inputs.

Gates (north star): bf16 path within 2e-2 relative of the reference's bf16-autocast results — loss, pred, latent,
per-block activations, every parameter gradient.
"""
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import TINY_CASES, oracle_cfg  # noqa: E402
from oracle import tae_oracle as O  # noqa: E402  (checker only)

BF16_TOL = 2e-2


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(kw):
    from tae_b200 import tae as T

    torch.manual_seed(0)
    m = T.TAE(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw)
    return m


@pytest.mark.parametrize("case", TINY_CASES)
def test_forward_backward_matches_golden_and_oracle(case, golden_meta, golden_tensors):
    rec = golden_meta[case]
    kw = rec["kwargs"]
    t = golden_tensors(case)
    model = build(kw)
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()}
    for n, (s, a) in rec["init_checksums"].items():  # seeded init == the reference's
        assert abs(float(sd_cpu[n].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), n
    model.cuda().train()
    x = t["input"].cuda()

    acts = {}
    hooks = [blk.register_forward_hook(lambda m, a, o, k=f"{pre}.{i}": acts.__setitem__(k, float(o.detach().float().norm())))
             for pre, blocks in (("blocks", model.blocks), ("decoder_blocks", model.decoder_blocks))
             for i, blk in enumerate(blocks)]
    loss, pred, latent = model(x, return_latent=True)
    for h in hooks:
        h.remove()
    assert loss.dtype == torch.float32 and loss.dim() == 0
    assert pred.dtype == torch.bfloat16 and latent.dtype == torch.bfloat16
    loss.backward()
    torch.cuda.synchronize()

    g = rec["bf16"]
    # (a) golden fixtures from the reference under bf16 autocast
    assert abs(float(loss) - g["loss"]) < BF16_TOL * abs(g["loss"])
    assert rel(pred.float().cpu(), t["bf16.pred"]) < BF16_TOL
    assert rel(latent.float().cpu(), t["bf16.latent"]) < BF16_TOL
    for k, v in g["block_out_norms"].items():
        assert abs(acts[k] - v) < BF16_TOL * v, k
    # and within bf16 noise of the reference's fp32 results
    assert abs(float(loss) - rec["fp32"]["loss"]) < BF16_TOL * rec["fp32"]["loss"]
    assert rel(pred.float().cpu(), t["fp32.pred"]) < BF16_TOL

    # (b) oracle on the GPU, same weights: every gradient
    sd = {k: v.cuda() for k, v in sd_cpu.items()}
    cfg = oracle_cfg(kw)
    lo, po, zo, go = O.forward_backward(sd, x, cfg, "bf16")
    assert abs(float(loss) - float(lo)) < 5e-3 * float(lo)
    assert rel(pred.float(), po.float()) < BF16_TOL
    assert rel(latent.float(), zo.float()) < BF16_TOL
    worst = ("", 0.0)
    for n, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, n
        e = rel(p.grad, go[n].float())
        if e > worst[1]:
            worst = (n, e)
        gn = float(p.grad.norm())
        assert abs(gn - g["grad_norm"][n]) < 3e-2 * g["grad_norm"][n] + 1e-7, (n, gn, g["grad_norm"][n])
    assert worst[1] < 3e-2, worst
    gnorm = float(torch.norm(torch.stack([p.grad.norm() for p in model.parameters()])))
    assert abs(gnorm - g["global_grad_norm"]) < BF16_TOL * g["global_grad_norm"]
    # 1-D gradients stored in full by the reference
    for n, p in model.named_parameters():
        key = f"bf16.grad.{n}"
        if key in t and float(t[key].norm()) > 1e-6:
            assert rel(p.grad.cpu(), t[key]) < 4e-2, n


@pytest.mark.parametrize("case", TINY_CASES)
def test_encoder_decoder_entry_points(case, golden_meta, golden_tensors):
    rec, t = golden_meta[case], golden_tensors(case)
    model = build(rec["kwargs"]).cuda().eval()
    x = t["input"].cuda()
    with torch.no_grad():
        z = model.forward_encoder(x)
        assert rel(z.float().cpu(), t["bf16.latent"]) < BF16_TOL
        assert torch.equal(model.encode(x), z)
        pred = model.forward_decoder(z)
        assert torch.equal(model.decode(z), pred)
        assert rel(pred.float().cpu(), t["bf16.pred"]) < BF16_TOL
        loss = model.forward_loss(x, pred)
        assert abs(float(loss) - rec["bf16"]["loss"]) < BF16_TOL * rec["bf16"]["loss"]
        loss2, pred2 = model(x)
        assert torch.equal(pred2, pred) and abs(float(loss2) - float(loss)) < 1e-6
        # reconstruction display path (train.py:190): unpatchify(pred) and its inverse
        img = model.unpatchify(pred)
        assert img.shape == x.shape
        assert torch.equal(model.patchify(img), pred)


def test_no_cpu_fallback():
    from tae_b200 import tae as T
    from tae_b200._lib import TaeError

    m = T.TAE(img_size=32, patch_size=8, embed_dim=128, depth=1, num_heads=2, decoder_embed_dim=128, decoder_depth=1,
              decoder_num_heads=2, vocab_size=16)
    with pytest.raises((TaeError, RuntimeError)):
        m(torch.randn(1, 3, 32, 32))  # CPU tensors must fail loudly, never fall back


def test_grad_accumulation_and_upstream_scale(golden_meta, golden_tensors):
    """loss/accum_iter over two micro-steps (train.py:145-148) == one step on the same data."""
    case = "tiny_p8_n16_hd32"
    rec, t = golden_meta[case], golden_tensors(case)
    model = build(rec["kwargs"]).cuda()
    x = t["input"].cuda()
    loss, _ = model(x)
    loss.backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    for _ in range(2):
        loss, _ = model(x)
        (loss / 2).backward()
    for n, p in model.named_parameters():
        assert rel(p.grad, ref[n]) < 5e-3, n
