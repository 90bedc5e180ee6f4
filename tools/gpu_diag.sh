#!/bin/bash
# fc1 (GELU epilogue) GEMM timing of several variant builds:  gpurun -- 'bash tools/gpu_diag.sh gelu8 tma8 oneout'
mkdir -p gpurun_out
for v in default "$@"; do
  [ "$v" = default ] && unset TAE_B200_LIB || export TAE_B200_LIB=tae_b200/libtae_b200.$v.so
  case $v in oneout|tmaoneout|default) ;; *)
    timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gelu or dynamic" --timeout 300 -p no:cacheprovider > gpurun_out/gpu_tests_$v.log 2>&1
    echo "$v gelu tests exit $?: $(tail -n 1 gpurun_out/gpu_tests_$v.log)";;
  esac
  timeout 200 python tools/gpu_probe.py --gemm-only --only fc1 > gpurun_out/probe_$v.log 2>&1
  echo "$v: $(grep -h 'fc1  fwd' gpurun_out/probe_$v.log)"
done
# ring attention forward (TAE_ATTN_FWD=ring): parity, then timing against the default persistent kernel
#   TAE_ATTN_FWD=ring python -m pytest tests -m gpu -q -k attention ; TAE_ATTN_FWD=ring python tools/gpu_probe.py --attn-only
