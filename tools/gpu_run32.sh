#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary32.txt; tail -n 5 gpurun_out/$name.log; }
run k_attn python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider -k "attention"
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from tae_b200 import ops
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
for (B,N,H,hd) in [(256,64,32,64),(256,64,12,64),(256,16,32,80)]:
    D=H*hd
    qkv=(torch.randn(B*N,3*D,device='cuda')*0.5).bfloat16(); dout=(torch.randn(B*N,D,device='cuda')*0.5).bfloat16()
    out,lse=ops.attention_fwd(qkv,B,N,H,hd)
    tf=t(lambda: ops.attention_fwd(qkv,B,N,H,hd)); tb=t(lambda: ops.attention_bwd(qkv,out,dout,lse,B,N,H,hd))
    by=qkv.numel()*2+out.numel()*2
    print(f"attention N={N} hd={hd} H={H}: fwd {tf*1e3:.1f} us ({by/tf/1e6:.0f} GB/s), bwd {tb*1e3:.1f} us ({(2*by)/tb/1e6:.0f} GB/s)")
PY
