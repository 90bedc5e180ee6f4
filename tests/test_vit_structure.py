"""CPU checks of the downstream ViTs (VITForRecognition / VITForSegmentation, reference tae.py:274-429) against the
fixtures generated from the unmodified reference (tests/golden/make_golden_vit.py): factory names, state_dict
keys / shapes / order of all 24 factories, and bit-identical seeded initialisation."""
import json
import os
from functools import partial

import pytest
import torch

from conftest import GOLDEN_DIR


@pytest.fixture(scope="module")
def vit_meta():
    with open(os.path.join(GOLDEN_DIR, "vit_golden_meta.json")) as f:
        return json.load(f)


def test_vit_factories_match_reference_state_dicts(vit_meta):
    from tae_b200 import tae as T

    table = vit_meta["factory_state_dicts"]
    assert sorted(table) == sorted(T.VIT_MODEL_NAMES) and len(table) == 24
    for name, ref in table.items():
        with torch.device("meta"):
            m = T.__dict__[name](num_classes=1000 if "recognition" in name else 21)
        mine = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert mine == ref, name


@pytest.mark.parametrize("case", ["vitrec_n16_hd32_c37", "vitrec_n256_hd64_c16", "vitseg_n16_p8_c5"])
def test_vit_seeded_init_matches_reference(case, vit_meta):
    from tae_b200 import tae as T

    rec = vit_meta[case]
    torch.manual_seed(0)
    m = getattr(T, rec["class"])(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **rec["kwargs"])
    assert [n for n, _ in m.named_parameters()] == rec["param_names"]
    for n, p in m.named_parameters():
        s, a = rec["init_checksums"][n]
        assert abs(float(p.detach().double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), n
        assert abs(float(p.detach().double().abs().sum()) - a) <= 1e-9 * max(1.0, a), n


def test_vit_no_cpu_fallback():
    from tae_b200 import _lib
    from tae_b200 import tae as T

    m = T.VITForRecognition(num_patches=4, vocab_size=16, decoder_embed_dim=128, decoder_depth=1, decoder_num_heads=2,
                            num_classes=8)
    with pytest.raises(_lib.TaeError):
        m(torch.randn(2, 4, 16))
