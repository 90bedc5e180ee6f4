#!/bin/bash
# A/B of two builds of the library on one box:  gpurun --timeout 900 -- 'bash tools/gpu_ab.sh scalar'
# (the variant library is built here first:  python -m tae_b200.build --variant scalar -D TAE_GELU_SCALAR)
v=${1:-scalar}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests exit $?"; tail -n 3 gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke.log
timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_default.log 2>&1
TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_$v.log 2>&1
grep -h "fc1  fwd" gpurun_out/probe_default.log gpurun_out/probe_$v.log
for i in 1 2; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default_$i.log 2>&1
  TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${v}_$i.log 2>&1
done
for f in gpurun_out/ab_default_*.log gpurun_out/ab_${v}_*.log; do
  python - "$f" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        rd = d.get("roofline_detail", {})
        print(sys.argv[1].split("/")[-1], round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms  gelu",
              round(rd.get("bf16_gelu", {}).get("tflops", 0)), "TF", d["clocks"]["sm_mhz"])
PY
done
