"""CPU tests of the drop-in boundary: model API surface, state_dict layout, host helpers, and the C ABI's symbols."""
import os
import re
from functools import partial

import pytest
import torch

from oracle import tae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_factories_state_dict_layout_matches_reference(golden_meta):
    """All 12 zero-arg factories: same names, key order and shapes as the reference's (tae.py:434-483)."""
    from tae_b200 import tae as T

    ref = golden_meta["factory_state_dicts"]
    assert sorted(T.MODEL_NAMES) == sorted(ref.keys()) and len(ref) == 12
    for name, spec in ref.items():
        assert name in T.__dict__, name  # drivers look models up via tae.__dict__[args.model]() (train.py:94)
        with torch.device("meta"):
            m = T.__dict__[name]()
        sd = m.state_dict()
        assert [[k, list(v.shape)] for k, v in sd.items()] == spec, name
        assert all(v.dtype == torch.float32 for v in sd.values())
        assert len(list(m.buffers())) == 0
        # the oracle's own table agrees
        assert [[k, list(s)] for k, s in O.state_dict_spec(O.zoo_config(name))] == spec, name


def test_model_api_surface():
    from tae_b200 import tae as T

    with torch.device("meta"):
        m = T.tae_patch32_vocab1024_px256()
    for attr in ("forward", "forward_encoder", "forward_decoder", "forward_loss", "patchify", "unpatchify", "encode",
                 "decode", "initialize_weights"):
        assert callable(getattr(m, attr)), attr
    assert m.patch_embed.num_patches == 64 and m.patch_embed.patch_size[0] == 32 and m.patch_embed.grid_size == (8, 8)
    assert m.pos_embed.shape == (1, 64, 2048) and m.decoder_pos_embed.shape == (1, 64, 2048)
    assert m.blocks[0].attn.num_heads == 32 and m.blocks[0].attn.head_dim == 64
    assert m.blocks[0].norm1.eps == 1e-6
    assert len(m.blocks) == 18 and len(m.decoder_blocks) == 18
    # constructor signature and defaults of tae.py:135-149
    import inspect

    sig = inspect.signature(T.TAE.__init__)
    assert list(sig.parameters)[1:] == ["img_size", "patch_size", "in_chans", "embed_dim", "vocab_size", "depth", "num_heads",
                                        "decoder_embed_dim", "decoder_depth", "decoder_num_heads", "mlp_ratio", "norm_layer"]
    assert sig.parameters["embed_dim"].default == 1024 and sig.parameters["decoder_embed_dim"].default == 512
    assert sig.parameters["depth"].default == 24 and sig.parameters["img_size"].default == 224


def test_no_cpu_fallback_raises():
    from tae_b200 import tae as T
    from tae_b200._lib import TaeError

    m = T.TAE(img_size=32, patch_size=8, embed_dim=128, depth=1, num_heads=2, decoder_embed_dim=128, decoder_depth=1,
              decoder_num_heads=2, vocab_size=16)
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(TaeError):
        m(torch.randn(1, 3, 32, 32))
    with pytest.raises(TaeError):
        m.forward_encoder(torch.randn(1, 3, 32, 32))
    with pytest.raises(TaeError):
        m.patchify(torch.randn(1, 3, 32, 32))


def test_patch_embed_asserts_like_reference():
    from tae_b200 import tae as T

    pe = T.PatchEmbed(img_size=32, patch_size=8, embed_dim=128)
    with pytest.raises(AssertionError):
        pe(torch.randn(1, 3, 40, 32))  # tae.py:48-49


def test_c_abi_exports_every_declared_symbol():
    from tae_b200 import _lib

    header = open(os.path.join(ROOT, "include", "tae_b200.h")).read()
    declared = set(re.findall(r"\b(tae_[a-z0-9_]+)\s*\(", header))
    declared -= {"tae_gemm_args"}
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tae_b200.h but not exported"
    assert set(_lib.PROTOTYPES) == declared
    assert lib.tae_version() == 1
    assert lib.tae_last_error_string() is not None
    # struct layout mirrors the header (field order)
    fields = re.search(r"typedef struct tae_gemm_args \{(.*?)\} tae_gemm_args;", header, re.S).group(1)
    cleaned = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    names = re.findall(r"(\w+)\s*[,;]", cleaned)
    assert [n for n, _ in _lib.GemmArgs._fields_] == [n for n in names if n], names


def test_weight_decay_groups_and_lr_schedule():
    from tae_b200 import misc
    from tae_b200 import tae as T

    with torch.device("meta"):
        m = T.tae_patch16_vocab16_px256()
    groups = misc.add_weight_decay(m, 0.05)
    assert groups[0]["weight_decay"] == 0.0 and groups[1]["weight_decay"] == 0.05
    names = {id(p): n for n, p in m.named_parameters()}
    no_decay = [names[id(p)] for p in groups[0]["params"]]
    decay = [names[id(p)] for p in groups[1]["params"]]
    nd_ref, d_ref = O.add_weight_decay_names([(n, tuple(p.shape)) for n, p in m.named_parameters()], 0.05)
    assert no_decay == nd_ref and decay == d_ref
    assert "pos_embed" in decay and "decoder_pos_embed" in decay and "dict_proj.weight" in decay
    assert all(n.endswith(".bias") or "norm" in n for n in no_decay)

    class Opt:
        param_groups = [{"lr": 0.0}, {"lr": 0.0, "lr_scale": 0.5}]

    opt = Opt()
    assert misc.adjust_learning_rate(opt, 1e-4, 1e-5, 10, 450000) == 1e-4
    assert opt.param_groups[0]["lr"] == 1e-4 and opt.param_groups[1]["lr"] == 5e-5
    assert misc.adjust_learning_rate(opt, 1e-4, 1e-5, 450000, 450000) == 1e-5
    assert O.adjust_learning_rate(1e-4, 1e-5, 449999, 450000) == 1e-4


def test_oracle_adamw_matches_torch_cpu():
    torch.manual_seed(3)
    p = torch.randn(257)
    g = torch.randn(257) * 0.1
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pt], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    po, m, v = p.clone(), torch.zeros(257), torch.zeros(257)
    for step in range(1, 5):
        pt.grad = g.clone()
        opt.step()
        po, m, v = O.adamw_step(po, g, m, v, step, 1e-3, 0.9, 0.95, 1e-8, 0.05)
    assert float((po - pt.detach()).abs().max()) < 1e-6


def test_shard_and_bucket_planning():
    from tae_b200.ddp import plan_buckets
    from tae_b200.engine import shard_for_rank

    assert [shard_for_rank(2048, r, 8) for r in range(8)] == [(256 * r, 256 * (r + 1)) for r in range(8)]
    assert shard_for_rank(10, 3, 4) == (9, 10) and shard_for_rank(10, 0, 4) == (0, 3)
    params = [torch.zeros(n) for n in (100, 30, 500, 64, 1)]
    offsets, off = {}, 0
    for p in params:
        offsets[id(p)] = off
        off += (p.numel() + 63) // 64 * 64
    buckets = plan_buckets(params, offsets, lambda p: p.numel(), 256)
    assert [len(b[2]) for b in buckets] == [3, 2]
    assert buckets[0][0] == 0 and buckets[0][1] == offsets[id(params[3])] and buckets[1][1] is None
    assert sum(len(b[2]) for b in buckets) == len(params)


def test_pos_embed_interpolation_and_load_model(tmp_path):
    from tae_b200 import misc
    from tae_b200 import tae as T

    kw = dict(img_size=32, patch_size=8, embed_dim=128, depth=1, num_heads=2, decoder_embed_dim=128, decoder_depth=1,
              decoder_num_heads=2, vocab_size=16, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    src, dst = T.TAE(**kw), T.TAE(**kw)
    import argparse

    ck = tmp_path / "ck.pth"
    torch.save({"model": src.state_dict(), "args": argparse.Namespace(model="x"), "iteration": 3}, ck)
    misc.load_model(str(ck), dst)  # the Namespace breaks weights_only loading in the reference on torch >= 2.6
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    big = {"pos_embed": torch.randn(1, 64, 128)}
    misc.interpolate_pos_embed(dst, big)
    assert big["pos_embed"].shape == (1, 16, 128)


def test_build_fingerprint_is_location_independent(tmp_path):
    """The GPU box runs a copy of the repo under another path: the prebuilt library must be recognised as current there
    (a rebuild by 8 torchrun ranks at once is how a half-written .so gets dlopen'ed)."""
    import importlib.util
    import shutil

    from tae_b200 import build as here

    root = tmp_path / "elsewhere"
    (root / "tae_b200").mkdir(parents=True)
    shutil.copytree(here.CSRC, root / "tae_b200" / "csrc", ignore=shutil.ignore_patterns("build"))
    shutil.copytree(here.INCLUDE_DIR, root / "include")
    shutil.copy(here.PKG_DIR / "build.py", root / "tae_b200" / "build.py")
    spec = importlib.util.spec_from_file_location("build_elsewhere", root / "tae_b200" / "build.py")
    there = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(there)
    assert there.INCLUDE_DIR != here.INCLUDE_DIR
    assert there._fingerprint() == here._fingerprint()
    stamp = here.BUILD_DIR / "fingerprint.txt"
    if here.LIB_PATH.exists() and stamp.exists():
        assert stamp.read_text() == here._fingerprint(), "in-tree library is stale: run python -m tae_b200.build"
