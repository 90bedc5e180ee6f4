"""ncu target: the 12 GEMMs of one patch16 transformer block (forward, dgrad, wgrad) at the bench shape M=65536."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import ops
from tae_b200._lib import *
M, D = 65536, 1024
bf = torch.bfloat16
r = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(bf)
x, wq, wp, w1, w2 = r(M, D), r(3 * D, D), r(D, D), r(4 * D, D), r(D, 4 * D)
bq, b1, b2 = torch.randn(3 * D, device="cuda"), torch.randn(4 * D, device="cuda"), torch.randn(D, device="cuda")
res = torch.randn(M, D, device="cuda")
part = torch.empty((M // 32, 4 * D), dtype=torch.float32, device="cuda")
for rep in range(2):
    qkv = ops.gemm(x, wq, epilogue=EPI_BF16, bias=bq)                        # 1 qkv fwd
    xm = ops.gemm(x, wp, epilogue=EPI_F32_RESID, bias=b2, resid=res)         # 2 proj fwd
    gp, a = ops.gemm(x, w1, epilogue=EPI_BF16_GELU, bias=b1)                 # 3 fc1 fwd
    y = ops.gemm(a, w2, epilogue=EPI_F32_RESID, bias=b2, resid=res)          # 4 fc2 fwd
    dw2 = ops.gemm(x, a, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)         # 5 fc2 wgrad
    dh = ops.gemm(x, w2, b_mn=True, epilogue=EPI_BF16_DGELU, aux=gp, colsum_partials=part)  # 6 fc2 dgrad
    dw1 = ops.gemm(dh, x, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)        # 7 fc1 wgrad
    dx = ops.gemm(dh, w1, b_mn=True, epilogue=EPI_BF16)                      # 8 fc1 dgrad
    dwp = ops.gemm(x, x, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)         # 9 proj wgrad
    da, dl2 = ops.gemm(x, wp, b_mn=True, epilogue=EPI_BF16_ROWDOT, aux=x, rowdot_tokens=256)  # 10 proj dgrad (+delta)
    dwq = ops.gemm(qkv, x, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)       # 11 qkv wgrad
    dl = ops.gemm(qkv, wq, b_mn=True, epilogue=EPI_BF16)                     # 12 qkv dgrad
torch.cuda.synchronize()
print("ok")
