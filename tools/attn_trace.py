"""Timeline of one CTA of the pipelined attention backward (clock64 stamps)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import ops, _lib
B, N, H, hd = 256, 256, 16, 64
D = H * hd
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
dout = (torch.randn(B * N, D, device="cuda") * 0.5).to(torch.bfloat16)
out, lse = ops.attention_fwd(qkv, B, N, H, hd)
for _ in range(3):
    ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd)
tr = torch.zeros(64, dtype=torch.int64, device="cuda")
L = _lib.load()
L.tae_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
L.tae_debug_set_attn_trace(tr.data_ptr())
ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd)
torch.cuda.synchronize()
L.tae_debug_set_attn_trace(0)
t = tr.tolist()
t0 = t[32 + 29]
print("persistent kernel, 6th item of CTA grid/2; t=0 at the element-wise warps' entry into the item")
print("E: delta ready", t[32] - t0)
for b in range(8):
    print(f"C blk{b}: bar_p seen {t[1+2*b]-t0:7d}  issued {t[2+2*b]-t0:7d}")
print("C: tail prefetch issued", t[17] - t0)
for b in range(8):
    print(f"E blk{b}: bar_s seen {t[32+1+3*b]-t0:7d}  computed {t[32+2+3*b]-t0:7d}  arrived {t[32+3+3*b]-t0:7d}")
print("E: final bar_g seen", t[32 + 25] - t0, " accumulators drained", t[32 + 26] - t0, " item done", t[32 + 27] - t0)
