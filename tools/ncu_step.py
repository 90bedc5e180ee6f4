"""ncu target: ONE full training step of the bench workload (tae_patch16_vocab256_px256, B=256) inside a
cudaProfilerStart/Stop range, after 2 warm-up steps.  Use with `ncu --profile-from-start off`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tae_b200 import engine, misc, ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
name = sys.argv[1] if len(sys.argv) > 1 else "tae_patch16_vocab256_px256"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
model = engine.build_model(name, dev).train()
opt = engine.build_optimizer(model, max_lr=1e-4, weight_decay=0.05)
scaler = misc.NativeScalerWithGradNormCount(compute_norm=False)
x = torch.randn(B, 3, 256, 256, device=dev)
for it in range(2):
    engine.train_step(model, opt, scaler, x, it, max_lr=1e-4)
torch.cuda.synchronize()
n0 = ops.launch_count()
torch.cuda.profiler.start()
loss = engine.train_step(model, opt, scaler, x, 2, max_lr=1e-4)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok loss", float(loss), "launches", ops.launch_count() - n0)
