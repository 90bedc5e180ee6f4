"""Golden fixtures for the downstream ViTs (VITForRecognition / VITForSegmentation, reference tae.py:274-429), generated
from the UNMODIFIED reference in the build container:

    python tests/golden/make_golden_vit.py        # needs /root/reference

Same recipe as make_golden.py: seeded init, seeded latent inputs, forward + backward in fp32 and under
torch.autocast('cpu', dtype=bfloat16); stored are the outputs in full, per-parameter gradient norms, all 1-D gradients
and the init checksums.  Losses: cross-entropy against seeded labels (recognition, as recognition/train_*.py) and
CE(out) + 0.5 CE(aux) (segmentation, as segmentation/train.py's criterion).
"""
import json
import os
import sys
from functools import partial

import torch
import torch.nn.functional as F

REF = os.environ.get("TAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (class, kwargs, batch)
    "vitrec_n16_hd32_c37": ("VITForRecognition", dict(num_patches=16, vocab_size=64, decoder_embed_dim=128, decoder_depth=2,
                                                       decoder_num_heads=4, mlp_ratio=4., num_classes=37), 3),
    "vitrec_n256_hd64_c16": ("VITForRecognition", dict(num_patches=256, vocab_size=16, decoder_embed_dim=128, decoder_depth=1,
                                                        decoder_num_heads=2, mlp_ratio=4., num_classes=16), 2),
    "vitseg_n16_p8_c5": ("VITForSegmentation", dict(num_patches=16, patch_size=8, vocab_size=64, decoder_embed_dim=128,
                                                     decoder_depth=4, decoder_num_heads=4, mlp_ratio=4., num_classes=5), 2),
}
MODEL_SEED, INPUT_SEED, LABEL_SEED = 0, 4321, 77


def loss_of(kind, out, labels):
    if kind == "VITForRecognition":
        return F.cross_entropy(out.float(), labels)
    return F.cross_entropy(out["out"].float(), labels) + 0.5 * F.cross_entropy(out["aux"].float(), labels)


def main():
    sys.path.insert(0, REF)
    import tae  # the reference, unmodified

    torch.set_num_threads(8)
    meta = {}
    for name, (cls, kw, batch) in CASES.items():
        torch.manual_seed(MODEL_SEED)
        model = getattr(tae, cls)(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw).train()
        z = torch.randn(batch, kw["num_patches"], kw["vocab_size"], generator=torch.Generator().manual_seed(INPUT_SEED))
        g = torch.Generator().manual_seed(LABEL_SEED)
        if cls == "VITForRecognition":
            labels = torch.randint(0, kw["num_classes"], (batch,), generator=g)
        else:
            S = int(kw["num_patches"] ** .5) * kw["patch_size"]
            labels = torch.randint(0, kw["num_classes"], (batch, S, S), generator=g)
        rec = {"class": cls, "kwargs": kw, "batch": batch, "torch": torch.__version__,
               "param_names": [n for n, _ in model.named_parameters()],
               "init_checksums": {n: [float(p.detach().double().sum()), float(p.detach().double().abs().sum())]
                                  for n, p in model.named_parameters()}}
        tensors = {"input": z, "labels": labels}
        for mode in ("fp32", "bf16"):
            model.zero_grad(set_to_none=True)
            if mode == "bf16":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    out = model(z)
            else:
                out = model(z)
            loss = loss_of(cls, out, labels)
            loss.backward()
            outs = {"out": out} if cls == "VITForRecognition" else out
            r = {"loss": float(loss), "out_dtype": {k: str(v.dtype) for k, v in outs.items()},
                 "grad_norm": {n: float(p.grad.float().norm()) for n, p in model.named_parameters()}}
            for k, v in outs.items():
                tensors[f"{mode}.{k}"] = v.detach().float().clone()
            for n, p in model.named_parameters():
                if p.grad.dim() == 1:
                    tensors[f"{mode}.grad.{n}"] = p.grad.detach().float().clone()
            rec[mode] = r
        meta[name] = rec
        torch.save(tensors, os.path.join(HERE, f"{name}.pt"))
        print(name, "fp32 loss", rec["fp32"]["loss"], "bf16 loss", rec["bf16"]["loss"])
    # state_dict tables of the 24 ViT factories (meta device)
    shapes = {}
    for fname in [n for n in dir(tae) if n.startswith("vit_")]:
        with torch.device("meta"):
            mm = getattr(tae, fname)(num_classes=1000 if "recognition" in fname else 21)
        shapes[fname] = [[k, list(v.shape)] for k, v in mm.state_dict().items()]
    meta["factory_state_dicts"] = shapes
    with open(os.path.join(HERE, "vit_golden_meta.json"), "w") as f:
        json.dump(meta, f)
    print("wrote vit_golden_meta.json")


if __name__ == "__main__":
    main()
