"""Pretty-print the last JSON line of a bench log."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("img/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
if "roofline" in d:
    r = d["roofline"]
    print("gemm TF", round(r["achieved"], 1), "frac", round(r["frac"], 3), "traffic/launch", r.get("traffic"), "alg bytes/launch",
          r.get("algorithmic_bytes_per_launch"))
    for k, v in d["roofline_detail"].items():
        val = v.get("tflops", v.get("GB/s", 0))
        frac = v.get("frac_of_hbm_peak", v.get("frac_of_bf16_peak"))
        print(f"  {k:16s} {val:9.1f} {'GB/s' if 'GB/s' in v else 'TF/s'}  {v['ms_per_step']:7.2f} ms/step" + (f"  frac {frac:.2f}" if frac else ""))
print(d.get("clocks"))
if "encode" in d:
    print("encode", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["encode"].items()})
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"])
for k, v in d.get("secondary", {}).items():
    print("secondary", k, round(v["value"], 1), "img/s", round(v["ms_per_step"], 2), "ms  e2e", round(v["e2e"]["value"], 1),
          "frac_sustained", round(v["frac_of_bf16_peak_sustained"], 3))
for k, v in d.get("gpu_reference", {}).items():
    print("gpu_reference", k, v if not isinstance(v, dict) else {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()})
if "ddp_check" in d:
    print("ddp_check", d["ddp_check"], d.get("ddp_check_detail"))
print("value_repeat", d.get("value_repeat"))
