"""ncu target: tcgen05 attention fwd + bwd at the bench shape (B=256, N=256, H=16, hd=64)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import ops
B, N, H, hd = 256, 256, 16, 64
D = H * hd
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
dout = (torch.randn(B * N, D, device="cuda") * 0.5).to(torch.bfloat16)
for rep in range(2):
    out, lse = ops.attention_fwd(qkv, B, N, H, hd)
    delta = (dout.float() * out.float()).view(B, N, H, hd).sum(-1).permute(0, 2, 1).contiguous()  # torch op: not profiled (-k regex:attn)
    dqkv = ops.attention_bwd(qkv, None, dout, lse, B, N, H, hd, delta=delta)                     # the production path
torch.cuda.synchronize()
print("ok")
