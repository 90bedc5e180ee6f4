"""ncu target: the short-grid kernels at the patch32 / patch64 bench shapes (B=256): N=64 attention (mma.sync path,
backward with the row-dot delta) and the row-group LayerNorm backward for D = 2048 / 2560."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import ops
bf = torch.bfloat16
B, N, H, hd = 256, 64, 32, 64
D = H * hd
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(bf)
dout = (torch.randn(B * N, D, device="cuda") * 0.5).to(bf)
ln = []
for rows, Dw in ((16384, 2048), (4096, 2560)):
    x = torch.randn(rows, Dw, device="cuda")
    w, b = torch.ones(Dw, device="cuda"), torch.zeros(Dw, device="cuda")
    _, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
    ln.append(((torch.randn(rows, Dw, device="cuda") * 0.5).to(bf), x, mean, rstd, w, torch.randn(rows, Dw, device="cuda")))
for rep in range(2):
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)                                  # attn_fwd_mma<64>
    delta = (dout.float() * out.float()).view(B, N, H, hd).sum(-1).permute(0, 2, 1).contiguous()
    ops.attention_bwd(qkv, None, dout, lse, B, N, H, hd, delta=delta)                # attn_bwd_mma64<true>
    for a in ln:
        ops.layernorm_bwd(*a)                                                        # ln_bwd_group_kernel<4,3> / <5,2>
torch.cuda.synchronize()
print("ok")
