#!/bin/bash
# Round-2 call A: full GPU suite on HEAD, the variant epilogues (parity + probe + bench A/B), ring attention A/B, and the
# complete default bench line (secondary configs, reference on the same GPU).
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; }
summ() {
  python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l)
        rd = d.get("roofline_detail", {})
        f = lambda k: round(rd.get(k, {}).get("tflops", 0))
        print(sys.argv[1].split("/")[-1], round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms | resid", f("f32_resid"),
              "rowdot", f("bf16_rowdot"), "gelu", f("bf16_gelu"), "TF | attn fwd/bwd ms",
              round(rd.get("attention_fwd", {}).get("ms_per_step", 0), 2), round(rd.get("attention_bwd", {}).get("ms_per_step", 0), 2),
              "| sm", d["clocks"]["sm_mhz"], "MHz")
PY
}
rm -f gpurun_out/grad_parity.txt
run gpu_tests python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider
ok=""
for v in rowdot resid both; do
  export TAE_B200_LIB=tae_b200/libtae_b200.$v.so
  timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bench_shapes.py tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/gpu_tests_$v.log 2>&1; rc=$?
  echo "$v gpu_tests exit $rc: $(tail -n 1 gpurun_out/gpu_tests_$v.log)"
  timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_$v.log 2>&1; grep -h "proj fwd\|proj dgrad\|fc2  fwd" gpurun_out/probe_$v.log
  [ $rc -eq 0 ] && ok="$ok $v"
  unset TAE_B200_LIB
done
timeout 200 python tools/gpu_probe.py --gemm-only > gpurun_out/probe_default.log 2>&1; grep -h "proj fwd\|proj dgrad\|fc2  fwd" gpurun_out/probe_default.log
AB="--steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-secondary --no-encode"
for i in 1 2; do
  timeout 300 python bench.py $AB > gpurun_out/ab_default_$i.log 2>&1; summ gpurun_out/ab_default_$i.log
  TAE_ATTN_FWD=ring timeout 300 python bench.py $AB > gpurun_out/ab_ring_$i.log 2>&1; summ gpurun_out/ab_ring_$i.log
  for v in $ok; do
    TAE_B200_LIB=tae_b200/libtae_b200.$v.so timeout 300 python bench.py $AB > gpurun_out/ab_${v}_$i.log 2>&1
    summ gpurun_out/ab_${v}_$i.log
  done
done
timeout 300 python bench.py $AB --graph > gpurun_out/ab_graph.log 2>&1; summ gpurun_out/ab_graph.log
/usr/bin/time -v timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench_full exit $?"
tail -c 6000 gpurun_out/bench_full.log
grep -E "Elapsed|Maximum resident" gpurun_out/bench_full.err
