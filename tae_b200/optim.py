"""Fused AdamW over flat parameter arenas (replaces `torch.optim.AdamW(param_groups, fused=True)`, train.py:109).

Each parameter group is flattened into five contiguous arenas — fp32 master parameters, fp32 gradients, exp_avg,
exp_avg_sq and a bf16 shadow copy of the parameters — laid out in REVERSE parameter order (the order gradients
become ready in backward), so that
  * one `tae_adamw_step` launch per group updates everything at HBM speed and refreshes the bf16 weights the
    GEMMs read (no separate cast pass, no 373 small launches);
  * the hand-written backward passes write weight gradients straight into the arena (`p._tae_grad`), with
    beta=0 on the first write after `zero_grad()` — no gradient memset and no autograd accumulate pass;
  * data-parallel buckets are contiguous slices of the gradient arena (tae_b200/ddp.py).
The object protocol matches what the reference's drivers use (train.py:108-116,146-148,165-168; util/misc.py:
400-412): `param_groups[i]["lr"]` is writable and read every step, `zero_grad()`, `state_dict()` /
`load_state_dict()` in torch.optim.AdamW's format (`step`, `exp_avg`, `exp_avg_sq` per parameter).
"""
from __future__ import annotations

import math

import torch

from . import ops

_ALIGN = 64  # elements; keeps every parameter 256-byte aligned in the fp32 arenas (128 B in the bf16 one)


class _Arena:
    def __init__(self, params, device):
        self.params = list(params)
        order = list(reversed(self.params))  # gradient-ready order
        self.offsets = {}
        off = 0
        for p in order:
            self.offsets[id(p)] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.order = order
        f32, dev = torch.float32, device
        self.p = torch.zeros(off, dtype=f32, device=dev)
        self.g = torch.zeros(off, dtype=f32, device=dev)
        self.m = torch.zeros(off, dtype=f32, device=dev)
        self.v = torch.zeros(off, dtype=f32, device=dev)
        self.pb = torch.zeros(off, dtype=torch.bfloat16, device=dev)
        # host-side bookkeeping that lets step() skip the per-parameter walk in the common case: how many parameters
        # had their gradient written by the hand-written backward since zero_grad().  A parameter that ALSO receives a
        # gradient through autograd's AccumulateGrad needs no flag: `p.grad` is the arena slice by then (or is folded in
        # by `_sink`), so the accumulation lands in the arena.  (No post-accumulate-grad hook: a hook keeps the
        # AccumulateGrad node — and the stream it was created on — alive, which breaks CUDA-graph capture.)
        self.n_direct = 0

    def view(self, arena, p):
        o = self.offsets[id(p)]
        return arena[o:o + p.numel()].view(p.shape)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0,
                 track_grad_norm=False, fused=True, capturable=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_scale = float(grad_scale)
        self.track_grad_norm = track_grad_norm
        self._arenas = []
        self._steps = []
        self._grad_sq = None
        self._hyper = None  # capturable mode: [n_groups, 9] derived scalars in device memory (see make_capturable)
        self._build()
        if capturable:
            self.make_capturable()

    # ------------------------------------------------------------------------------------------
    def _build(self):
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._arenas.append(None)
                self._steps.append(0)
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise RuntimeError("FusedAdamW needs CUDA parameters (move the model to the GPU first); no CPU fallback")
            for p in ps:
                if p.dtype != torch.float32 or p.device != dev:
                    raise RuntimeError("FusedAdamW: parameters must be fp32 and on one device")
            ar = _Arena(ps, dev)
            with torch.no_grad():
                for p in ps:
                    ar.view(ar.p, p).copy_(p)
                    old_grad = p.grad
                    p.data = ar.view(ar.p, p)
                    g = ar.view(ar.g, p)
                    if old_grad is not None:
                        g.copy_(old_grad)
                    p.grad = g if old_grad is not None else None
                    p._tae_grad = g
                    p._tae_arena = ar
                    p._tae_dirty = 1 if old_grad is not None else 0
                    ar.n_direct += p._tae_dirty
                    p._tae_bf16 = ar.view(ar.pb, p)
                    p._tae_bf16_version = p._version
                    p._tae_bf16_ptr = p.data_ptr()
                ops.cast_bf16(ar.p, ar.pb)
            self._arenas.append(ar)
            self._steps.append(0)
        dev = next((a.p.device for a in self._arenas if a is not None), None)
        if dev is not None:
            self._grad_sq = torch.zeros(1, dtype=torch.float32, device=dev)

    @property
    def arenas(self):
        return [a for a in self._arenas if a is not None]

    # ------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        """Gradients live in the arena permanently; 'zeroing' arms beta=0 for the next backward's first write and drops
        `p.grad` (as torch's set_to_none=True does).  The hand-written backward re-points `p.grad` at the arena slice
        when it writes it; a gradient that arrives through autograd's own AccumulateGrad (a torch-native op using the
        parameter) therefore lands in a fresh tensor, never on stale arena contents, and is folded in by `_sink` /
        `step()`."""
        for ar in self.arenas:
            for p in ar.params:
                p._tae_dirty = 0
                p.grad = None
            ar.n_direct = 0

    # ------------------------------------------------------------------------------------------
    # CUDA-graph support: scalars in device memory
    @property
    def capturable(self) -> bool:
        return self._hyper is not None

    def make_capturable(self):
        """From now on step() reads lr / bias corrections / weight decay from a device buffer that `prepare_step()`
        refreshes, so a captured graph of the step can be replayed while `param_groups[i]["lr"]` keeps changing
        (util/misc.py:400-412).  Results are bit-identical to the host-scalar path (same derived floats)."""
        if self._hyper is None:
            dev = next(a.p.device for a in self._arenas if a is not None)
            self._hyper = torch.zeros((len(self.param_groups), ops.ADAMW_HYPER_FLOATS), dtype=torch.float32, device=dev)
        return self

    def prepare_step(self):
        """Advance the step counters and upload this step's scalars (one small H2D copy from pageable memory, which the
        driver stages before returning — the host may run ahead).  Called by step() itself outside graph capture, and by
        the owner of a captured graph before every replay."""
        rows = []
        for gi, (group, ar) in enumerate(zip(self.param_groups, self._arenas)):
            if ar is None:
                rows.append([0.0] * ops.ADAMW_HYPER_FLOATS)
                continue
            self._steps[gi] += 1
            b1, b2 = group["betas"]
            rows.append(ops.adamw_hyper(lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"],
                                        weight_decay=group["weight_decay"], step=self._steps[gi], grad_scale=self.grad_scale))
        self._hyper.copy_(torch.tensor(rows, dtype=torch.float32), non_blocking=True)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._grad_sq is not None and self.track_grad_norm:
            self._grad_sq.zero_()
        if self._hyper is not None and not torch.cuda.is_current_stream_capturing():
            self.prepare_step()
        for gi, (group, ar) in enumerate(zip(self.param_groups, self._arenas)):
            if ar is None:
                continue
            if ar.n_direct != len(ar.params):
                # slow path (never taken by a plain TAE step): fold in gradients that autograd accumulated itself, and
                # give parameters without any gradient a ZERO gradient.  torch.optim.AdamW skips such parameters
                # entirely (no weight decay, no moment decay); one launch over a flat arena cannot, so here they decay
                # as if their gradient were zero — documented difference, irrelevant for TAE (every parameter is used).
                for p in ar.params:
                    g, t = p.grad, p._tae_grad
                    foreign = g is not None and g.data_ptr() != t.data_ptr()
                    if p._tae_dirty == 0:
                        if foreign:
                            t.copy_(g)
                        else:
                            t.zero_()
                    elif foreign:
                        t.add_(g)
                    if foreign:
                        p.grad = t
                        p._tae_dirty = 1
            gsq = self._grad_sq if self.track_grad_norm else None
            if self._hyper is not None:
                ops.adamw_step_dev(ar.p, ar.g, ar.m, ar.v, ar.pb, self._hyper[gi], grad_sq_sum=gsq)
                continue
            self._steps[gi] += 1
            b1, b2 = group["betas"]
            ops.adamw_step(ar.p, ar.g, ar.m, ar.v, ar.pb, lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"],
                           weight_decay=group["weight_decay"], step=self._steps[gi], grad_scale=self.grad_scale,
                           grad_sq_sum=gsq)
        return loss

    def grad_norm(self) -> torch.Tensor:
        """L2 norm of all (scaled) gradients seen by the last step() (needs track_grad_norm=True); device tensor."""
        return self._grad_sq.sqrt().squeeze(0)

    # ------------------------------------------------------------------------------------------
    # torch.optim.AdamW-compatible checkpoint format
    def state_dict(self):
        state, groups, idx = {}, [], 0
        for gi, (group, ar) in enumerate(zip(self.param_groups, self._arenas)):
            ids = []
            for p in group["params"]:
                ids.append(idx)
                if ar is not None and id(p) in ar.offsets and self._steps[gi] > 0:
                    state[idx] = {"step": torch.tensor(float(self._steps[gi])),
                                  "exp_avg": ar.view(ar.m, p).clone(), "exp_avg_sq": ar.view(ar.v, p).clone()}
                idx += 1
            g = {k: v for k, v in group.items() if k != "params"}
            g["params"] = ids
            groups.append(g)
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, sd):
        idx = 0
        for gi, (group, ar, saved) in enumerate(zip(self.param_groups, self._arenas, sd["param_groups"])):
            for k, v in saved.items():
                if k != "params":
                    group[k] = v
            step = 0
            for p in group["params"]:
                st = sd["state"].get(idx, sd["state"].get(str(idx)))
                if st is not None and ar is not None and id(p) in ar.offsets:
                    ar.view(ar.m, p).copy_(st["exp_avg"])
                    ar.view(ar.v, p).copy_(st["exp_avg_sq"])
                    step = max(step, int(float(st["step"])))
                idx += 1
            self._steps[gi] = step
