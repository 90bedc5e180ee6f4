#!/bin/bash
# A/B of variant builds of the library against the default one, on one box:
#   python -m tae_b200.build --variant tma -D TAE_GELU_TMA_EPI=1        (here, before the call)
#   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh tma [more variants]'
# Each variant first has to pass the GPU parity tests (TAE_B200_LIB selects the library), then is timed.
mkdir -p gpurun_out
summ() {
  python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        rd = d.get("roofline_detail", {})
        print(sys.argv[1].split("/")[-1], round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms  gelu",
              round(rd.get("bf16_gelu", {}).get("tflops", 0)), "TF  sm", d["clocks"]["sm_mhz"], "MHz")
PY
}
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "default gpu_tests exit $?"; tail -n 1 gpurun_out/gpu_tests.log
timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_default.log 2>&1; grep -h "fc1  fwd" gpurun_out/probe_default.log
ok=""
for v in "$@"; do
  export TAE_B200_LIB=tae_b200/libtae_b200.$v.so
  timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/gpu_tests_$v.log 2>&1; rc=$?
  echo "$v gpu_tests exit $rc"; tail -n 1 gpurun_out/gpu_tests_$v.log
  timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_$v.log 2>&1; grep -h "fc1  fwd" gpurun_out/probe_$v.log
  [ $rc -eq 0 ] && ok="$ok $v"
  unset TAE_B200_LIB
done
for i in 1 2; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary --no-encode > gpurun_out/ab_default_$i.log 2>&1; summ gpurun_out/ab_default_$i.log
  for v in $ok; do
    TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary --no-encode > gpurun_out/ab_${v}_$i.log 2>&1
    summ gpurun_out/ab_${v}_$i.log
  done
done
