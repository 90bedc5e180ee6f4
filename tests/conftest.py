import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN_DIR, "golden_meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_tensors():
    import torch

    cache = {}
    with open(os.path.join(GOLDEN_DIR, "golden_meta.json")) as f:
        meta = json.load(f)

    def load(name):
        if name not in cache:
            t = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), map_location="cpu")
            if "input" not in t:  # compact fixtures: the input is regenerated from its recorded seed (make_golden.py)
                rec = meta[name]
                S = rec["kwargs"]["img_size"]
                t["input"] = torch.randn(rec["batch"], 3, S, S, generator=torch.Generator().manual_seed(rec["seeds"]["input"]))
            cache[name] = {k: (v.float() if v.dtype == torch.bfloat16 else v) for k, v in t.items()}
        return cache[name]

    return load


TINY_CASES = ["tiny_p8_n64_hd64", "tiny_p8_n16_hd32", "tiny_p16_n4_hd80"]
# + the real patch16 width / grid / head size with one block per side, generated from the unmodified reference as well
GOLDEN_CASES = TINY_CASES + ["mid_p16_n256_hd64"]


def oracle_cfg(kw):
    from oracle import tae_oracle as O

    return O.TAEConfig(kw["img_size"], kw["patch_size"], kw["embed_dim"], kw["vocab_size"], kw["depth"], kw["num_heads"],
                       kw["decoder_embed_dim"], kw["decoder_depth"], kw["decoder_num_heads"], kw["mlp_ratio"])


def report(tag, rows):
    """Worst tensors per model, printed (pytest -s / -rA) and appended to gpurun_out/grad_parity.txt on the GPU box."""
    import os

    rows = sorted(rows, key=lambda r: -r[1])[:5]
    line = tag + ": " + ", ".join(f"{n} {e:.2e}" for n, e in rows)
    print(line)
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/grad_parity.txt", "a") as f:
            f.write(line + "\n")
    return rows[0] if rows else ("", 0.0)
