#!/bin/bash
# Evidence run: ncu launch list of one bench step + ncu --set full of every kernel family (second repetition of each tool script).
mkdir -p gpurun_out
R=${1:-r2}
STEP="python tools/ncu_step.py"
$STEP > gpurun_out/ncu_plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$R.csv $STEP > gpurun_out/ncu_step.log 2>&1
echo "launch list rc $?"
python tools/ncu_gemm.py > gpurun_out/ncu_plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm -s 12 -c 12 -f -o gpurun_out/gemm_$R python tools/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc $?"
python tools/ncu_attn.py > gpurun_out/ncu_plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn -s 2 -c 2 -f -o gpurun_out/attn_$R python tools/ncu_attn.py > gpurun_out/ncu_attn.log 2>&1
echo "attn rc $?"
python tools/ncu_misc.py > gpurun_out/ncu_plain_misc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'ln_|colsum|mse|im2col|adamw' -s 9 -c 9 -f -o gpurun_out/misc_$R python tools/ncu_misc.py > gpurun_out/ncu_misc.log 2>&1
echo "misc rc $?"
for f in ncu_step ncu_gemm ncu_attn ncu_misc; do tail -n 3 gpurun_out/$f.log; done
python tools/ncu_short.py > gpurun_out/ncu_plain_short.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_mma|attn_bwd_mma64|ln_bwd_group' -s 4 -c 4 -f -o gpurun_out/short_$R python tools/ncu_short.py > gpurun_out/ncu_short.log 2>&1
echo "short rc $?"; tail -n 3 gpurun_out/ncu_short.log
