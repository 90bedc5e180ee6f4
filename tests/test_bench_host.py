"""CPU tests of bench.py's host logic: metric labels, the algorithmic FLOP model, the gate on the ncu traffic
evidence, the staged reference modules, and the JSON contract of the reference arm (which runs without a GPU)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_metric_label_follows_the_model():
    assert bench.metric_name("tae_patch16_vocab256_px256") == "train images/sec at px256 patch16"
    assert bench.metric_name("tae_patch128_vocab16384_px256") == "train images/sec at px256 patch128"
    assert bench.metric_name("tae_patch64_vocab4096_px256", "encode") == "encode images/sec at px256 patch64"
    assert {k: v[0] for k, v in bench.CONFIGS.items()} == {
        2: "tae_patch16_vocab256_px256", 3: "tae_patch32_vocab1024_px256", 4: "tae_patch128_vocab16384_px256",
        5: "tae_patch64_vocab4096_px256"}
    assert bench.CONFIGS[5][1] == "encode" and all(bench.CONFIGS[c][1] == "train" for c in (2, 3, 4))


@pytest.mark.parametrize("N,D,L,V,gf", [(256, 1024, 15, 16, 606.04), (256, 1024, 15, 256, 606.80), (64, 2048, 18, 1024, 705.05),
                                        (16, 2560, 21, 4096, 324.47), (4, 2560, 22, 16384, 90.12)])
def test_train_flops_match_the_survey_table(N, D, L, V, gf):
    """F_train = 3 F_fwd - 2*196608*D with attention backward credited 2x forward (GF per image)."""
    assert abs(bench.train_flops_per_image(N, D, L, V, 256) / 1e9 - gf) < 0.01


def test_traffic_evidence_is_bound_to_the_gemm_sources():
    """roofline.traffic is only quoted from a profiles/*_traffic.json captured from the very GEMM sources of the checkout."""
    fp = bench.gemm_sources_fingerprint()
    h = hashlib.sha256()
    for name in ("gemm_sm100.cu", "sm100.cuh", "common.cuh"):
        with open(os.path.join(ROOT, "tae_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    assert fp == h.hexdigest()[:16]
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
        tj = json.load(f)
    assert "gemm_bf16_tcgen05" in tj["kernels"] and tj["kernels"]["gemm_bf16_tcgen05"]["dram_bytes_per_launch"] > 0
    # committed evidence must describe the committed kernel: re-capture (tools/gpu_profile.sh) after touching the GEMM
    assert tj["gemm_sources_fingerprint"] == fp, "profiles/r2_traffic.json is stale for the current GEMM sources"


def test_staged_reference_is_byte_identical():
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "MANIFEST.sha256")):
        pytest.skip("oracle/_ref not staged (oracle/make_ref.sh needs /root/reference)")
    for line in open(os.path.join(ref, "MANIFEST.sha256")):
        digest, name = line.split()
        with open(os.path.join(ref, name), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, name
        src = os.path.join("/root/reference", name)
        if os.path.exists(src):  # build container: the copy is the reference, unmodified
            with open(src, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, name
    mods = bench.load_reference_modules()
    assert mods is not None and hasattr(mods[0], "tae_patch16_vocab256_px256") and hasattr(mods[1], "add_weight_decay")


@pytest.mark.timeout(600)
def test_reference_arm_json_contract():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the contract's keys; `steps` = steps timed."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "1", "--model", "tae_patch16_vocab16_px256"], capture_output=True, text=True, timeout=580)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "train images/sec at px256 patch16" and line["unit"] == "images/s"
    assert line["steps"] == 1 and line["warmup"] == 0 and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("tae_patch16_vocab16_px256")


@pytest.mark.timeout(600)
def test_cpu_config1_reproduces_the_pinned_loss():
    """BASELINE.json configs[0]: the reference's CPU-runnable case, as bench.py reports it beside the GPU numbers."""
    r = bench.cpu_config1(repeats=1)
    assert r["model"] == "tae_patch16_vocab16_px256" and r["batch"] == 2 and r["kind"] in ("reference", "port")
    assert r["pinned_loss_ok"] and abs(r["loss"] - 2.184759855) < 1e-4 * 2.184759855
    for k in ("fwd_loss", "forward_encoder", "fwd_bwd"):
        assert r[k]["best_s"] > 0 and r[k]["images_per_s"] > 0
