"""Generate the golden fixtures in this directory from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only mount of eminorhan/tae)

The reference's `tae.py` is imported as-is (sys.path), instantiated with small constructor arguments under a fixed
seed, and run on seeded inputs: forward (+ per-block activations), backward, and one AdamW step with the parameter
groups of the reference's `util/misc.add_weight_decay`.  Runs in fp32 and under
`torch.autocast('cpu', dtype=torch.bfloat16)`.  What is stored (see tests/golden/README.md):
  * config, seeds, per-parameter checksums of the seeded init (so init parity can be checked where the reference
    is absent);
  * loss, pred, latent (full tensors), per-block activation norms;
  * per-parameter gradient norm and the dot product of the gradient with a seeded probe tensor; all 1-D gradients in full;
  * per-parameter norm of the AdamW update after one step (wd 0.05 on the decay group, lr 1e-3).
The fixtures are the pin for oracle/tae_oracle.py (tests/test_oracle_golden.py) and, on the GPU box where
/root/reference does not exist, for the CUDA path (tests/test_gpu_model.py).
"""
import json
import os
import sys

import torch

REF = os.environ.get("TAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (constructor kwargs, batch)
    "tiny_p8_n64_hd64": (dict(img_size=64, patch_size=8, in_chans=3, embed_dim=128, vocab_size=16, depth=2, num_heads=2,
                              decoder_embed_dim=128, decoder_depth=2, decoder_num_heads=2, mlp_ratio=4.), 2),
    "tiny_p8_n16_hd32": (dict(img_size=32, patch_size=8, in_chans=3, embed_dim=128, vocab_size=64, depth=1, num_heads=4,
                              decoder_embed_dim=128, decoder_depth=1, decoder_num_heads=4, mlp_ratio=4.), 3),
    "tiny_p16_n4_hd80": (dict(img_size=32, patch_size=16, in_chans=3, embed_dim=640, vocab_size=256, depth=1, num_heads=8,
                              decoder_embed_dim=640, decoder_depth=1, decoder_num_heads=8, mlp_ratio=4.), 4),
    # the real patch16 width, grid and head size (D=1024, N=256, hd=64: the tcgen05 attention and CTA-pair GEMM paths) with
    # one block per side; the input is regenerated from its seed and bf16 tensors are stored as bf16 to keep the file small
    "mid_p16_n256_hd64": (dict(img_size=256, patch_size=16, in_chans=3, embed_dim=1024, vocab_size=256, depth=1, num_heads=16,
                               decoder_embed_dim=1024, decoder_depth=1, decoder_num_heads=16, mlp_ratio=4.), 2),
}
COMPACT = {"mid_p16_n256_hd64"}
MODEL_SEED, INPUT_SEED, PROBE_SEED = 0, 1234, 99


def probe_like(t, k):
    g = torch.Generator().manual_seed(PROBE_SEED + k)
    return torch.randn(t.shape, generator=g)


def run_case(tae, misc, name, kwargs, batch):
    from functools import partial

    torch.manual_seed(MODEL_SEED)
    model = tae.TAE(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kwargs)
    model.train()
    x = torch.randn(batch, 3, kwargs["img_size"], kwargs["img_size"], generator=torch.Generator().manual_seed(INPUT_SEED))
    out = {"name": name, "kwargs": kwargs, "batch": batch, "seeds": dict(model=MODEL_SEED, input=INPUT_SEED, probe=PROBE_SEED),
           "torch": torch.__version__}
    names = [n for n, _ in model.named_parameters()]
    out["param_names"] = names
    out["init_checksums"] = {n: [float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
                             for n, p in model.named_parameters()}

    tensors = {}
    for mode in ("fp32", "bf16"):
        acts = {}
        hooks = []
        for prefix, blocks in (("blocks", model.blocks), ("decoder_blocks", model.decoder_blocks)):
            for i, blk in enumerate(blocks):
                hooks.append(blk.register_forward_hook(
                    lambda m, a, o, key=f"{prefix}.{i}": acts.__setitem__(key, float(o.detach().float().norm()))))
        model.zero_grad(set_to_none=True)
        if mode == "bf16":
            with torch.autocast("cpu", dtype=torch.bfloat16):
                latent = model.forward_encoder(x)
                pred = model.forward_decoder(latent)
                loss = model.forward_loss(x, pred)
        else:
            latent = model.forward_encoder(x)
            pred = model.forward_decoder(latent)
            loss = model.forward_loss(x, pred)
        loss.backward()
        for h in hooks:
            h.remove()
        rec = {"loss": float(loss), "block_out_norms": acts,
               "pred_dtype": str(pred.dtype), "latent_dtype": str(latent.dtype), "loss_dtype": str(loss.dtype)}
        keep = (lambda t: t.detach().clone()) if (name in COMPACT and mode == "bf16") else (lambda t: t.detach().float().clone())
        tensors[f"{mode}.pred"] = keep(pred)
        tensors[f"{mode}.latent"] = keep(latent)
        gn, gp = {}, {}
        for k, (n, p) in enumerate(model.named_parameters()):
            g = p.grad.detach().float()
            gn[n] = float(g.norm())
            gp[n] = float((g.double() * probe_like(g, k).double()).sum())
            if g.dim() == 1:
                tensors[f"{mode}.grad.{n}"] = g.clone()
        rec["grad_norm"] = gn
        rec["grad_probe"] = gp
        rec["global_grad_norm"] = float(misc.get_grad_norm_(list(model.parameters())))
        out[mode] = rec

    # one AdamW step exactly as train.py:108-109 builds it (fp32 grads from the last, bf16-autocast, backward are
    # replaced by the fp32 ones: rerun fp32 backward)
    model.zero_grad(set_to_none=True)
    loss, _ = model(x)
    loss.backward()
    groups = misc.add_weight_decay(model, 0.05, bias_wd=False)
    out["no_decay_names"] = [n for n, p in model.named_parameters() if any(p is q for q in groups[0]["params"])]
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt = torch.optim.AdamW(groups, lr=1e-3, betas=(0.9, 0.95))
    opt.step()
    out["adamw"] = {"lr": 1e-3, "betas": [0.9, 0.95], "weight_decay": 0.05,
                    "update_norm": {n: float((p.detach() - before[n]).norm()) for n, p in model.named_parameters()},
                    "param_norm_after": {n: float(p.detach().norm()) for n, p in model.named_parameters()}}
    if name not in COMPACT:
        tensors["input"] = x  # compact cases: torch.randn(batch, 3, S, S, generator=manual_seed(seeds.input)) regenerates it
    return out, tensors


def main():
    sys.path.insert(0, REF)
    import tae  # the reference, unmodified
    from util import misc

    torch.set_num_threads(8)
    meta = {}
    only = [a for a in sys.argv[1:] if a in CASES]  # `make_golden.py CASE...` adds cases to the existing fixtures
    if only:
        with open(os.path.join(HERE, "golden_meta.json")) as f:
            meta = json.load(f)
    for name, (kwargs, batch) in CASES.items():
        if only and name not in only:
            continue
        rec, tensors = run_case(tae, misc, name, kwargs, batch)
        meta[name] = rec
        torch.save(tensors, os.path.join(HERE, f"{name}.pt"))
        print(name, "fp32 loss", rec["fp32"]["loss"], "bf16 loss", rec["bf16"]["loss"])
    if only:
        with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
            json.dump(meta, f)
        return

    # seed-pinned scalars of the real config #1, re-measured here from the reference itself
    torch.manual_seed(0)
    m = tae.tae_patch16_vocab16_px256()
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(1234))
    loss, pred = m(x)
    loss.backward()
    with torch.no_grad():
        lat = m.forward_encoder(x)
    meta["tae_patch16_vocab16_px256_b2"] = {
        "loss": float(loss), "pred_abs_mean": float(pred.detach().abs().mean()), "latent_abs_mean": float(lat.abs().mean()),
        "global_grad_norm": float(misc.get_grad_norm_(list(m.parameters()))),
        "init_checksums": {n: [float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
                           for n, p in list(m.named_parameters())[:8] + list(m.named_parameters())[-6:]},
        "state_dict_keys": list(m.state_dict().keys()),
    }
    print("patch16_vocab16 b2:", {k: v for k, v in meta["tae_patch16_vocab16_px256_b2"].items() if isinstance(v, float)})

    # state_dict key/shape tables of all 12 factories (meta device: no memory)
    shapes = {}
    for fname in [n for n in dir(tae) if n.startswith("tae_patch")]:
        with torch.device("meta"):
            mm = getattr(tae, fname)()
        shapes[fname] = [[k, list(v.shape)] for k, v in mm.state_dict().items()]
    meta["factory_state_dicts"] = shapes
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f)
    print("wrote", os.path.join(HERE, "golden_meta.json"))


if __name__ == "__main__":
    main()
