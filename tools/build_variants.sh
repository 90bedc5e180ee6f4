#!/bin/bash
# Variant builds of the library for A/B runs (tools/gpu_ab.sh NAME...).  Each line: name and the macros that differ from
# the default build (see the switch list at the top of tae_b200/csrc/gemm_sm100.cu).
set -e
build() { name=$1; shift; python -m tae_b200.build --variant "$name" "$@" > /dev/null && echo "built tae_b200/libtae_b200.$name.so ($*)"; }
case "${1:-all}" in
  rowdot|all) build rowdot -D TAE_ROWDOT_TMA_EPI=1 ;;&
  resid|all)  build resid  -D TAE_RESID_TMA_EPI=1 ;;&
  both|all)   build both   -D TAE_ROWDOT_TMA_EPI=1 -D TAE_RESID_TMA_EPI=1 ;;&
  generic)    build generic -D TAE_GELU_TMA_EPI=0 -D TAE_GELU_EW=16 -D TAE_DGELU_TMA_EPI=0 ;;  # the pre-row-layout epilogues
esac
