#!/bin/bash
mkdir -p gpurun_out
for m in tae_patch32_vocab1024_px256 tae_patch64_vocab4096_px256 tae_patch128_vocab16384_px256; do
  timeout 600 python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline --no-encode > gpurun_out/bench_$m.log 2>&1; echo "$m exit $?"
  tail -n 1 gpurun_out/bench_$m.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('img/s',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'model TF',round(d.get('model_tflops_per_gpu',0),1),'gemm TF',round(d['roofline']['achieved'],1),'launches',d['gpu_launches'])
print({k:(round(v['tflops']),round(v['ms_per_step'],2),v['launches_per_step']) for k,v in d['roofline_detail'].items()})
print(d['clocks'])"
done
