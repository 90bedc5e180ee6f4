"""Turn ncu outputs brought back in gpurun_out/ into the markdown summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches_train_step.md
    python tools/summarize_ncu.py full gpurun_out/gemm_r1.ncu-rep      > profiles/r1_ncu_gemm.md
"""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "time us"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])


def short(name: str) -> str:
    name = name.replace("void ", "").replace("tae::", "")
    cut = name.find("(CUtensorMap")
    if cut < 0:
        cut = name.find("(const")
    if cut < 0:
        cut = name.find("(")
    return name[:cut] if cut > 0 else name[:80]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        if r[iu] == "ns":
            v /= 1e3
        elif r[iu] == "ms":
            v *= 1e3
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    print(f"{n} launches, {total / 1e3:.2f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES)\n")
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / total:.1f}% |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), lab) for m, lab in METRICS.items() if m in hdr]
    ik = hdr.index("Kernel Name")
    print("| # | kernel | " + " | ".join(f"{lab} [{units[i]}]" if units[i] else lab for i, lab in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for n, r in enumerate(rows[2:]):
        vals = []
        for i, _ in cols:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {n} | `{short(r[ik])}` | " + " | ".join(vals) + " |")


def traffic(*paths):
    """Average DRAM bytes (read + write) per launch and kernel family -> JSON for bench.py's `roofline.traffic`."""
    import json

    fam = OrderedDict()
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        it = hdr.index("gpu__time_duration.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
        for r in rows[2:]:
            name = short(r[ik])
            key = "gemm_bf16_tcgen05" if "gemm_bf16_tcgen05" in name else name.split("<")[0]
            b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            d = fam.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
            d["launches"] += 1
            d["dram_bytes"] += b
            d["time_us"] += float(r[it]) * tscale[units[it]]
    res = {k: {"launches_profiled": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
               "avg_time_us": v["time_us"] / v["launches"]} for k, v in fam.items()}
    import os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import gemm_sources_fingerprint  # bench.py only quotes the GEMM traffic for the sources it was captured from

    print(json.dumps({"gemm_sources_fingerprint": gemm_sources_fingerprint(),
                      "source": "ncu --set full (tools/gpu_profile.sh); dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                "averaged over the profiled launches of each kernel family at the bench shapes",
                      "kernels": res}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
