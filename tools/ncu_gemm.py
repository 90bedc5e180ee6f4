"""ncu target: one launch of each GEMM epilogue variant at the patch16 block shapes (M=16384)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import ops
from tae_b200._lib import *
M, D = 16384, 1024
bf = torch.bfloat16
r = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(bf)
x, w1, w2 = r(M, D), r(4 * D, D), r(D, 4 * D)
b1, b2 = torch.randn(4 * D, device="cuda"), torch.randn(D, device="cuda")
res = torch.randn(M, D, device="cuda")
for rep in range(2):
    h, a = ops.gemm(x, w1, epilogue=EPI_BF16_GELU, bias=b1)                 # fc1 fwd
    y = ops.gemm(a, w2, epilogue=EPI_F32_RESID, bias=b2, resid=res)         # fc2 fwd
    dh = ops.gemm(x, w2, b_mn=True, epilogue=EPI_BF16_DGELU, aux=h)         # fc2 dgrad
    dx = ops.gemm(dh, w1, b_mn=True, epilogue=EPI_BF16)                     # fc1 dgrad
    dw = ops.gemm(dh, x, a_mn=True, b_mn=True, epilogue=EPI_F32_ACC)        # fc1 wgrad
    p = ops.gemm(x, r(D, D), epilogue=EPI_F32_RESID, bias=b2, resid=res)    # proj fwd
torch.cuda.synchronize()
print("ok")
