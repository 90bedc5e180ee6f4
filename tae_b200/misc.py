"""Host-side runtime helpers on the hot path, mirroring the reference's util/misc.py (same names and argument
meaning) so a driver written against the reference reads the same here.

  init_distributed_mode         util/misc.py:212-242   (torchrun / SLURM / single process; NCCL)
  NativeScalerWithGradNormCount util/misc.py:245-271   (loss.backward -> [unscale] -> grad norm -> step)
  get_grad_norm_                util/misc.py:274-286
  add_weight_decay              util/misc.py:364-379
  adjust_learning_rate          util/misc.py:400-412
  load_model / interpolate_pos_embed  util/misc.py:307-350 (state_dict contract; checkpoint fix-ups of SURVEY §8f.4)
"""
from __future__ import annotations

import os
import sys
from math import inf

import torch
import torch.distributed as dist


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def save_on_master(*args, **kwargs):
    if is_main_process():
        torch.save(*args, **kwargs)


def init_distributed_mode(args):
    """One process per GPU.  Rank discovery: torchrun env (RANK/LOCAL_RANK/WORLD_SIZE), else SLURM_PROCID, else a
    single process.  (The reference reads SLURM_PROCID even in the torchrun branch, util/misc.py:217, which breaks
    plain torchrun — fixed here without changing the attributes it sets: args.rank/gpu/world_size.)"""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.world_size = int(os.environ["WORLD_SIZE"])
        args.rank = int(os.environ["RANK"])
        args.gpu = int(os.environ.get("LOCAL_RANK", args.rank % max(torch.cuda.device_count(), 1)))
    elif "SLURM_PROCID" in os.environ:
        args.world_size = int(os.environ.get("WORLD_SIZE", os.environ.get("SLURM_NTASKS", "1")))
        args.rank = int(os.environ["SLURM_PROCID"])
        args.gpu = args.rank % torch.cuda.device_count()
    elif torch.cuda.is_available():
        args.rank, args.gpu, args.world_size = 0, 0, 1
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
    else:
        print("Does not support training without GPU.")
        sys.exit(1)
    torch.cuda.set_device(args.gpu)
    dist.init_process_group(backend="nccl", init_method=getattr(args, "dist_url", "env://"),
                            world_size=args.world_size, rank=args.rank, device_id=torch.device("cuda", args.gpu))
    dist.barrier()


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    parameters = [p for p in parameters if p.grad is not None]
    norm_type = float(norm_type)
    if len(parameters) == 0:
        return torch.tensor(0.)
    device = parameters[0].grad.device
    if norm_type == inf:
        return max(p.grad.detach().abs().max().to(device) for p in parameters)
    return torch.norm(torch.stack([torch.norm(p.grad.detach(), norm_type).to(device) for p in parameters]), norm_type)


class NativeScalerWithGradNormCount:
    """Same call protocol as the reference's AMP scaler wrapper (util/misc.py:245-271).

    The reference trains in fp16 with `torch.amp.GradScaler`; the B200 path computes in bf16 (fp32-range exponent),
    where loss scaling is a no-op, so the scale is fixed at 1 and no unscale / inf-check passes over the gradients
    are needed.  With tae_b200.optim.FusedAdamW the gradient norm comes out of the optimizer's own pass
    (`track_grad_norm=True`) instead of 373 separate reductions; with any other optimizer it falls back to
    get_grad_norm_.  `compute_norm=False` skips it entirely (train.py:146 discards the value)."""

    state_dict_key = "amp_scaler"

    def __init__(self, compute_norm: bool = True):
        self.compute_norm = compute_norm

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        loss.backward(create_graph=create_graph)
        norm = None
        if update_grad:
            if clip_grad is not None:
                assert parameters is not None
                norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
                optimizer.step()
            elif self.compute_norm and getattr(optimizer, "track_grad_norm", False):
                optimizer.step()
                norm = optimizer.grad_norm()
            else:
                if self.compute_norm and parameters is not None:
                    norm = get_grad_norm_(parameters)
                optimizer.step()
        return norm

    def state_dict(self):
        return {"scale": 1.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 0}

    def load_state_dict(self, state_dict):
        pass


def add_weight_decay(model, weight_decay=1e-5, skip_list=(), bias_wd=False):
    """Two param groups exactly as util/misc.py:364-379: [no_decay (1-D or *.bias), decay (the rest)]."""
    decay, no_decay = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if (not bias_wd) and len(param.shape) == 1 or name.endswith(".bias") or name in skip_list:
            no_decay.append(param)
        else:
            decay.append(param)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


def adjust_learning_rate(optimizer, max_lr, min_lr, it, switch_it):
    """Step schedule (util/misc.py:400-412): max_lr until switch_it, min_lr after."""
    lr = max_lr if it < switch_it else min_lr
    for param_group in optimizer.param_groups:
        param_group["lr"] = lr * param_group["lr_scale"] if "lr_scale" in param_group else lr
    return lr


def interpolate_pos_embed(model, checkpoint_model):
    """Bicubic resize of a checkpoint's pos_embed when the patch grid differs (util/misc.py:326-350)."""
    if "pos_embed" not in checkpoint_model:
        return
    pe = checkpoint_model["pos_embed"]
    dim = pe.shape[-1]
    num_patches = model.patch_embed.num_patches
    extra = model.pos_embed.shape[-2] - num_patches
    old, new = int((pe.shape[-2] - extra) ** 0.5), int(num_patches ** 0.5)
    if old != new:
        tok = pe[:, extra:].reshape(-1, old, old, dim).permute(0, 3, 1, 2)
        tok = torch.nn.functional.interpolate(tok, size=(new, new), mode="bicubic", align_corners=False)
        checkpoint_model["pos_embed"] = torch.cat((pe[:, :extra], tok.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)


def load_model(ckpt, model_without_ddp, optimizer=None, loss_scaler=None, optim_resume=False):
    """util/misc.py:307-323.  Checkpoints written by the reference pickle an argparse.Namespace under 'args', which
    `weights_only=True` rejects on torch >= 2.6; the Namespace class is allow-listed for that load."""
    if not ckpt:
        return
    if ckpt.startswith("https"):
        checkpoint = torch.hub.load_state_dict_from_url(ckpt, map_location="cpu", check_hash=True)
    else:
        import argparse

        with torch.serialization.safe_globals([argparse.Namespace]):
            checkpoint = torch.load(ckpt, weights_only=True, map_location="cpu")
    interpolate_pos_embed(model_without_ddp, checkpoint["model"])
    model_without_ddp.load_state_dict(checkpoint["model"], strict=False)
    print(f"Resumed checkpoint {ckpt}")
    if "optimizer" in checkpoint and optim_resume and optimizer is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
        if "scaler" in checkpoint and loss_scaler is not None:
            loss_scaler.load_state_dict(checkpoint["scaler"])
        print("With optim & sched!")
