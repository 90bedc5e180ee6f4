#!/bin/bash
# Everything that was written at the end of round 1 without GPU time, in one call:
#   bash tools/build_variants.sh all && python -m tae_b200.build --variant poly2 -D TAE_ATTN_EXP2_POLY=2   (here)
#   gpurun --timeout 900 -- 'bash tools/gpu_pending.sh'
# 1. ring attention forward (TAE_ATTN_FWD=ring): parity incl. the two-key-halves cases, then timing against the default
# 2. row-dot / residual epilogue variants: parity (incl. the ragged variant cases) + GEMM probe + one bench A/B each
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?: $(tail -n 1 gpurun_out/$name.log | cut -c1-160)"; }
run attn_default python tools/gpu_probe.py --attn-only
TAE_ATTN_FWD=ring run attn_ring_tests python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -k "attention or model" --timeout 300 -p no:cacheprovider
TAE_ATTN_FWD=ring run attn_ring python tools/gpu_probe.py --attn-only
if [ -f tae_b200/libtae_b200.poly2.so ]; then
  TAE_ATTN_FWD=ring TAE_B200_LIB=tae_b200/libtae_b200.poly2.so run attn_ring_poly2_tests python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" --timeout 300 -p no:cacheprovider
  TAE_ATTN_FWD=ring TAE_B200_LIB=tae_b200/libtae_b200.poly2.so run attn_ring_poly2 python tools/gpu_probe.py --attn-only
fi
grep -h "attention fwd" gpurun_out/attn_default.log gpurun_out/attn_ring.log gpurun_out/attn_ring_poly2.log 2>/dev/null
vars=""
for v in rowdot resid both; do [ -f tae_b200/libtae_b200.$v.so ] && vars="$vars $v"; done
[ -n "$vars" ] && bash tools/gpu_ab.sh $vars
