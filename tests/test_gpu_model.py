"""Model-level parity on the GPU: the CUDA path against (a) the golden fixtures generated from the unmodified
reference (tests/golden/make_golden.py) and (b) the oracle on the same weights and inputs.

Gates (north star): bf16 path within 2e-2 relative of the reference's bf16-autocast results — loss, pred, latent,
per-block activations, every parameter gradient.
"""
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN_CASES, oracle_cfg, report  # noqa: E402
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


@pytest.mark.parametrize("case", GOLDEN_CASES)
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
    vs_oracle, vs_norm, vs_ref = [], [], []
    for n, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, n
        vs_oracle.append((n, rel(p.grad, go[n].float())))
        gn = float(p.grad.norm())
        vs_norm.append((n, abs(gn - g["grad_norm"][n]) / (g["grad_norm"][n] + 1e-7)))
        key = f"bf16.grad.{n}"  # 1-D gradients stored in full by the reference
        if key in t and float(t[key].norm()) > 1e-6:
            vs_ref.append((n, rel(p.grad.cpu(), t[key])))
    worst = [report(f"{case} grad vs oracle", vs_oracle), report(f"{case} grad norm vs reference", vs_norm),
             report(f"{case} 1-D grad vs reference", vs_ref)]
    # north star: every gradient within 2e-2 of the reference's bf16 path
    assert all(w[1] < BF16_TOL for w in worst), worst
    gnorm = float(torch.norm(torch.stack([p.grad.norm() for p in model.parameters()])))
    assert abs(gnorm - g["global_grad_norm"]) < BF16_TOL * g["global_grad_norm"]


@pytest.mark.parametrize("case", GOLDEN_CASES)
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
    # the no_grad forward skips what only a backward would need (gelu'(h)); the values are those of the training forward
    model.train()
    _, pred_t, z_t = model(x, return_latent=True)
    assert torch.equal(pred_t, pred) and torch.equal(z_t, z)
    model.eval()
    with torch.no_grad():
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


# ----------------------------------------------------------------------------------------------------
# The real model zoo at reduced batch: every code path of BASELINE.json's configs (N = 256/64/16/4 tokens,
# head_dim 64/80, latent widths 256..16384, split-K weight gradients, tcgen05 and generic attention) against the
# oracle on the same device, same weights, same inputs.
# ----------------------------------------------------------------------------------------------------
FULL_CASES = [("tae_patch16_vocab256_px256", 4), ("tae_patch32_vocab1024_px256", 4), ("tae_patch64_vocab4096_px256", 3),
              ("tae_patch128_vocab16384_px256", 2), ("tae_patch16_vocab16_px256", 2)]


@pytest.mark.parametrize("name,batch", FULL_CASES)
def test_full_size_models_match_oracle(name, batch):
    from tae_b200 import tae as T

    torch.manual_seed(0)
    with torch.device("cuda"):
        model = T.__dict__[name]()
    model.train()
    cfg = O.zoo_config(name)
    x = torch.randn(batch, 3, 256, 256, generator=torch.Generator().manual_seed(1234)).cuda()
    loss, pred, latent = model(x, return_latent=True)
    loss.backward()
    torch.cuda.synchronize()
    assert pred.shape == (batch, cfg.num_patches, 3 * cfg.patch_size ** 2) and latent.shape == (batch, cfg.num_patches, cfg.vocab_size)

    # oracle forward (no autograd graph needed for the activations)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    with torch.no_grad():
        lo, po, zo = O.forward(sd, x, cfg, "bf16")
    assert abs(float(loss) - float(lo)) < 5e-3 * float(lo), (float(loss), float(lo))
    assert rel(pred.float(), po.float()) < BF16_TOL
    assert rel(latent.float(), zo.float()) < BF16_TOL
    del po, zo

    # gradients: oracle autograd on the same weights, compared per tensor, then freed block by block
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    lo2, _, _ = O.forward(leaves, x, cfg, "bf16")
    lo2.backward()
    errs = []
    for n, p in model.named_parameters():
        g, go = p.grad, leaves[n].grad
        assert g is not None and go is not None, n
        errs.append((n, rel(g, go.float())))
    worst = report(f"{name} B={batch} grad vs oracle", errs)
    assert worst[1] < BF16_TOL, (worst, sum(e >= BF16_TOL for _, e in errs))
    gn = float(torch.norm(torch.stack([p.grad.norm() for p in model.parameters()])))
    gno = float(torch.norm(torch.stack([v.grad.float().norm() for v in leaves.values()])))
    assert abs(gn - gno) < BF16_TOL * gno


def test_encode_sharding_matches_unsharded():
    """encode.py path, batch-sharded (engine.shard_for_rank): concatenated shard outputs == the unsharded batch."""
    from tae_b200 import engine
    from tae_b200 import tae as T

    torch.manual_seed(0)
    with torch.device("cuda"):
        model = T.tae_patch64_vocab4096_px256()
    model.eval()
    x = torch.randn(8, 3, 256, 256, generator=torch.Generator().manual_seed(7)).cuda()
    full = engine.encode_batch(model, x)
    parts = []
    for r in range(4):
        lo, hi = engine.shard_for_rank(8, r, 4)
        parts.append(engine.encode_batch(model, x[lo:hi].contiguous()))
    assert torch.equal(torch.cat(parts, 0), full)
    assert full.shape == (8, 16, 4096) and full.dtype == torch.bfloat16


def test_fused_adamw_training_matches_torch_adamw():
    """FusedAdamW (direct-gradient arenas, bf16 shadows) vs torch.optim.AdamW on the same model, 3 steps."""
    from tae_b200 import engine, misc
    from tae_b200 import tae as T

    kw = dict(img_size=64, patch_size=8, in_chans=3, embed_dim=128, vocab_size=16, depth=2, num_heads=2,
              decoder_embed_dim=128, decoder_depth=2, decoder_num_heads=2, mlp_ratio=4.)
    x = torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(3)).cuda()
    torch.manual_seed(0)
    m1 = T.TAE(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw).cuda()
    torch.manual_seed(0)
    m2 = T.TAE(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw).cuda()
    o1 = engine.build_optimizer(m1, max_lr=1e-3, weight_decay=0.05, track_grad_norm=True)
    o2 = torch.optim.AdamW(misc.add_weight_decay(m2, 0.05), lr=1e-3, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    for it in range(3):
        l1, _ = m1(x)
        norm = scaler(l1, o1, parameters=m1.parameters())
        o1.zero_grad()
        l2, _ = m2(x)
        l2.backward()
        ref_norm = misc.get_grad_norm_(m2.parameters())
        o2.step()
        o2.zero_grad(set_to_none=True)
        assert abs(float(l1) - float(l2)) < 2e-3 * float(l2), it
        assert abs(float(norm) - float(ref_norm)) < 1e-2 * float(ref_norm), it
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if n.endswith("attn.qkv.bias"):
            # the key bias has an exactly-zero true gradient (softmax shift invariance): Adam normalises pure
            # rounding noise to +-lr there, so only the q and v thirds are comparable
            D = p1.numel() // 3
            p1, p2 = torch.cat([p1[:D], p1[2 * D:]]), torch.cat([p2[:D], p2[2 * D:]])
        assert rel(p1, p2) < 5e-3, n
    sd = o1.state_dict()
    ref = o2.state_dict()
    assert len(sd["state"]) == len(ref["state"]) and [g["params"] for g in sd["param_groups"]] == [g["params"] for g in ref["param_groups"]]
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


# ----------------------------------------------------------------------------------------------------
# fp32 ("no autocast") mode: 1e-4 relative gate of the north star, against the reference's own fp32 results
# ----------------------------------------------------------------------------------------------------
FP32_TOL = 1e-4


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_fp32_mode_matches_reference_fp32(case, golden_meta, golden_tensors):
    rec = golden_meta[case]
    kw = rec["kwargs"]
    t = golden_tensors(case)
    model = build(kw)
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()}
    model.cuda().train().set_precision("fp32")
    assert model.precision == "fp32"
    x = t["input"].cuda()
    acts = {}
    hooks = [blk.register_forward_hook(lambda m, a, o, k=f"{pre}.{i}": acts.__setitem__(k, float(o.detach().float().norm())))
             for pre, blocks in (("blocks", model.blocks), ("decoder_blocks", model.decoder_blocks))
             for i, blk in enumerate(blocks)]
    loss, pred, latent = model(x, return_latent=True)
    for h in hooks:
        h.remove()
    assert loss.dtype == pred.dtype == latent.dtype == torch.float32  # the reference's dtypes without autocast
    loss.backward()
    torch.cuda.synchronize()

    g = rec["fp32"]
    # (a) golden fixtures generated by the unmodified reference in fp32
    assert abs(float(loss) - g["loss"]) < FP32_TOL * abs(g["loss"])
    assert rel(pred.cpu(), t["fp32.pred"]) < FP32_TOL
    assert rel(latent.cpu(), t["fp32.latent"]) < FP32_TOL
    for k, v in g["block_out_norms"].items():  # per-layer activations
        assert abs(acts[k] - v) < FP32_TOL * v, k
    for n, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, n
        gn = float(p.grad.norm())
        assert abs(gn - g["grad_norm"][n]) < 2e-4 * g["grad_norm"][n] + 1e-9, (n, gn, g["grad_norm"][n])
        key = f"fp32.grad.{n}"
        if key in t and float(t[key].norm()) > 1e-6:
            assert rel(p.grad.cpu(), t[key]) < 2e-4, n
    gnorm = float(torch.norm(torch.stack([p.grad.norm() for p in model.parameters()])))
    assert abs(gnorm - g["global_grad_norm"]) < FP32_TOL * g["global_grad_norm"]

    # (b) the oracle in fp32 on the same weights: EVERY gradient tensor
    sd = {k: v.cuda() for k, v in sd_cpu.items()}
    lo, po, zo, go = O.forward_backward(sd, x, oracle_cfg(kw), "fp32")
    assert abs(float(loss) - float(lo)) < FP32_TOL * float(lo)
    assert rel(pred, po) < FP32_TOL and rel(latent, zo) < FP32_TOL
    worst = max(((rel(p.grad, go[n]), n) for n, p in model.named_parameters()))
    assert worst[0] < 2e-4, worst

    # switching back restores the bf16 path on the same module
    model.set_precision("bf16")
    model.zero_grad(set_to_none=True)
    l2, p2 = model(x)
    assert p2.dtype == torch.bfloat16 and abs(float(l2) - g["loss"]) < BF16_TOL * g["loss"]


def test_fp32_mode_encode_decode_no_grad(golden_meta, golden_tensors):
    case = "tiny_p8_n16_hd32"
    rec, t = golden_meta[case], golden_tensors(case)
    model = build(rec["kwargs"]).cuda().eval().set_precision("fp32")
    with torch.no_grad():
        z = model.forward_encoder(t["input"].cuda())
        y = model.forward_decoder(z)
    assert rel(z.cpu(), t["fp32.latent"]) < FP32_TOL and rel(y.cpu(), t["fp32.pred"]) < FP32_TOL


def test_latent_writer_matches_synchronous_reference_layout(tmp_path, golden_meta, golden_tensors):
    """encode.py:80-100 restated: the asynchronous writer produces the reference's file layout with identical contents."""
    from tae_b200 import engine

    case = "tiny_p8_n16_hd32"
    rec, t = golden_meta[case], golden_tensors(case)
    model = build(rec["kwargs"]).cuda().eval()
    xs = [torch.randn(5, 3, 32, 32, generator=torch.Generator().manual_seed(s)).cuda() for s in range(7)]
    ys = [torch.arange(5) + 5 * s for s in range(7)]
    ref_lat, ref_tgt = [], []
    w = engine.LatentWriter(str(tmp_path / "lat.pth"), depth=2)
    w2 = engine.LatentWriter(str(tmp_path / "shards.pth"), depth=3, shard_rows=10)
    for x, y in zip(xs, ys):
        z = engine.encode_batch(model, x)
        w.put(z, y)
        w2.put(z, y)
        ref_lat.append(z.cpu())  # what the reference does (synchronous)
        ref_tgt.append(y)
    files = w.close()
    assert files == [str(tmp_path / "lat.pth")]
    d = torch.load(files[0])
    assert set(d) == {"latents", "targets"}
    assert torch.equal(d["latents"], torch.cat(ref_lat)) and torch.equal(d["targets"], torch.cat(ref_tgt))
    parts = w2.close()
    assert len(parts) == 4  # 35 rows in parts of >= 10
    cat = torch.cat([torch.load(p)["latents"] for p in parts])
    assert torch.equal(cat, torch.cat(ref_lat))
    assert w.bytes_copied == sum(z.numel() * z.element_size() for z in ref_lat)
