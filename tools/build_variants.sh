#!/bin/bash
# Variant builds of the library for A/B runs (tools/gpu_ab.sh NAME...).  Each line: name and the macros that differ from
# the default build (see the switch list at the top of tae_b200/csrc/gemm_sm100.cu).  Round 2 decided every pending variant
# (profiles/r2_ab_results.md); what is left is the pre-row-layout baseline of the GELU / GELU' epilogues.
set -e
build() { name=$1; shift; python -m tae_b200.build --variant "$name" "$@" > /dev/null && echo "built tae_b200/libtae_b200.$name.so ($*)"; }
case "${1:-generic}" in
  generic) build generic -D TAE_GELU_TMA_EPI=0 -D TAE_GELU_EW=16 -D TAE_DGELU_TMA_EPI=0 ;;
  *) echo "usage: $0 [generic]   (other variants: python -m tae_b200.build --variant NAME -D MACRO=VALUE)"; exit 1 ;;
esac
