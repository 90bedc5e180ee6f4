"""ncu target: ONE encode batch (forward_encoder under no_grad, encode.py:80-88) of tae_patch64_vocab4096_px256, B=256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import engine, ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
name = sys.argv[1] if len(sys.argv) > 1 else "tae_patch64_vocab4096_px256"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
model = engine.build_model(name, dev).eval()
x = torch.randn(B, 3, 256, 256, device=dev)
for it in range(3):
    z = engine.encode_batch(model, x)
torch.cuda.synchronize()
n0 = ops.launch_count()
torch.cuda.profiler.start()
z = engine.encode_batch(model, x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(z.shape), "launches", ops.launch_count() - n0)
