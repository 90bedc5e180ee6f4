"""ncu target: the HBM-bound kernels at the bench shape (M=65536, D=1024; 380M-parameter AdamW arena)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import ops
M, D = 65536, 1024
bf = torch.bfloat16
xf = torch.randn(M, D, device="cuda")
w, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
dy = (torch.randn(M, D, device="cuda") * 0.5).to(bf)
dres = torch.randn(M, D, device="cuda")
big = (torch.randn(M, 3 * D, device="cuda") * 0.5).to(bf)
imgs = torch.randn(256, 3, 256, 256, device="cuda")
pred = (torch.randn(256, 256, 768, device="cuda") * 0.5).to(bf)
n = 380_000_000
p, g, m, v = (torch.randn(n, device="cuda") * 0.01 for _ in range(4))
v = v.abs()
pb = torch.empty(n, dtype=bf, device="cuda")
for rep in range(2):
    y, mean, rstd = ops.layernorm_fwd(xf, w, b, 1e-6)
    ops.layernorm_bwd(dy, xf, mean, rstd, w, dres)
    ops.colsum(big)
    ops.mse_loss(pred, imgs, 16, want_grad=True)
    ops.im2col(imgs, 16)
    ops.adamw_step(p, g, m, v, pb, lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0, step=3)
torch.cuda.synchronize()
print("ok")
