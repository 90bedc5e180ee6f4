#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('img/s',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'gemm TF',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3))
print({k:(round(v['tflops']),round(v['ms_per_step'],2)) for k,v in d['roofline_detail'].items()})
print(d['clocks'])"
