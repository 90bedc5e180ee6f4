"""Batch-sharded data parallelism: bucketed gradient all-reduce over NCCL (NVLink 5 / NVSwitch), overlapped with
backward.  Replaces `DistributedDataParallel(model, device_ids=[gpu])` (train.py:102, evaluate.py:73).

Design (one process per GPU):
  * parameters are broadcast from rank 0 once, as a handful of large arena broadcasts;
  * gradient buckets are CONTIGUOUS SLICES of FusedAdamW's gradient arena, which is laid out in gradient-ready
    order — no flatten / unflatten copies, a bucket is reduced in place;
  * the hand-written backward passes call `p._tae_ready(p)` the moment a parameter's gradient has been written;
    when a bucket's last parameter reports in, the compute stream records an event and the bucket's
    `all_reduce(AVG)` is enqueued on a dedicated high-priority communication stream, so NVLink traffic overlaps
    the remaining backward GEMMs (the reference's 25 MiB c10d buckets; default here 64 MiB, NVSwitch-sized);
  * an end-of-backward callback makes the compute stream wait for the communication stream, so
    `optimizer.step()` sees reduced gradients without a host synchronisation.
`no_sync()` skips the reduction for gradient-accumulation micro-steps (the reference reduces on every micro-step,
train.py:137-148, which is numerically equivalent and strictly more traffic).

The bucket bookkeeping is device-agnostic and is exercised with the gloo backend on CPU in tests/test_ddp_gloo.py.
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.autograd import Variable


class _Bucket:
    __slots__ = ("flat", "params", "pending", "work", "seen")

    def __init__(self, flat, params):
        self.flat, self.params, self.work, self.seen = flat, params, None, set()
        self.pending = len(params)

    def rearm(self):
        self.pending, self.work = len(self.params), None
        self.seen.clear()


def plan_buckets(order, offsets, numel_of, bucket_elems, tail_elems=0):
    """Split an arena (parameters `order` at element `offsets`) into contiguous buckets of ~bucket_elems elements.
    Returns [(start, end, [params...])] covering the arena in order.

    `tail_elems` > 0 cuts a separate LAST bucket of at most that many elements: the gradients that become ready last
    (PatchEmbed, the pos-embeds) then travel in a short all-reduce of their own — it is the one nothing can overlap —
    instead of at the end of a full-size bucket that cannot leave before them."""
    tail_from = len(order)
    if tail_elems > 0:
        acc = 0
        while tail_from > 1:
            p = order[tail_from - 1]
            end = offsets[id(order[tail_from])] if tail_from < len(order) else offsets[id(p)] + numel_of(p)
            acc = end - offsets[id(order[tail_from - 1])] + acc
            if acc > tail_elems:
                break
            tail_from -= 1
    buckets, cur, start = [], [], None
    for i, p in enumerate(order):
        o = offsets[id(p)]
        if start is None:
            start = o
        cur.append(p)
        end = offsets[id(order[i + 1])] if i + 1 < len(order) else None
        size = (end if end is not None else o + numel_of(p)) - start
        if size >= bucket_elems or end is None or i + 1 == tail_from:
            buckets.append((start, end, cur))
            cur, start = [], None
    return buckets


class DistributedDataParallel(nn.Module):
    def __init__(self, module: nn.Module, optimizer=None, bucket_mb: float = 64.0, process_group=None,
                 device_ids=None, broadcast: bool = True, tail_mb: float = 8.0, overlap: bool = True):
        super().__init__()
        self.module = module
        # overlap=False: the buckets are all-reduced back to back at the END of backward instead of while it runs — the
        # whole exchange is then exposed, but no NCCL kernel holds SMs (and HBM bandwidth) under the backward GEMMs
        self.overlap = overlap
        self.process_group = process_group
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.tail_elems = int(tail_mb * (1 << 20) / 4)
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._buckets = None
        self._require_sync = True
        self._callback_queued = False
        self._comm_stream = None
        self._launched = []
        self._optimizer = optimizer
        self._needs_broadcast = broadcast and self.world_size > 1
        if self.world_size > 1 and next(module.parameters()).is_cuda:
            # NCCL's all-reduce CTAs will hold SMs while the weight-gradient GEMMs run: a persistent kernel with a static
            # work list doubles its run time when one of its CTAs cannot become resident, so switch the library to
            # dynamic work lists (patch32_v1024 at 8 GPUs: 9666 -> 10400 img/s; ~0.5 % slower on an unshared GPU)
            from . import ops

            ops.set_dynamic_scheduling(True)
        if optimizer is not None:
            self.attach(optimizer)

    # -- setup ---------------------------------------------------------------------------------------
    def attach(self, optimizer):
        """Bind to a FusedAdamW (its arenas define the buckets).  Called automatically on the first forward when the
        parameters already carry arena views."""
        arenas = optimizer.arenas
        self._optimizer = optimizer
        self._buckets = []
        for ar in arenas:
            for start, end, params in plan_buckets(ar.order, ar.offsets, lambda p: p.numel(), self.bucket_elems,
                                                   self.tail_elems):
                flat = ar.g[start:end if end is not None else ar.numel]
                b = _Bucket(flat, params)
                self._buckets.append(b)
                for p in params:
                    p._tae_bucket = b
                    p._tae_ready = self._on_ready
        if self._needs_broadcast:
            for ar in arenas:
                dist.broadcast(ar.p, src=0, group=self.process_group)
                ar.pb.copy_(ar.p)  # refresh the bf16 shadows from the broadcast masters
            self._needs_broadcast = False
        dev = arenas[0].p.device if arenas else None
        if dev is not None and dev.type == "cuda":
            self._comm_stream = torch.cuda.Stream(device=dev, priority=-1)

    def _lazy_attach(self):
        if self._buckets is not None:
            return
        opt = self._optimizer
        if opt is None:
            raise RuntimeError("tae_b200.DistributedDataParallel needs the FusedAdamW that owns the gradient arenas: "
                               "pass optimizer=... or call ddp.attach(optimizer) before the first training step")
        self.attach(opt)

    # -- forward -------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        if torch.is_grad_enabled() and self.world_size > 1:
            self._lazy_attach()
            if self._launched:
                raise RuntimeError("tae_b200.DistributedDataParallel: a previous backward left reduced buckets un-finalised")
        elif self._needs_broadcast and self._optimizer is None:
            # inference-only replica: broadcast the plain parameters once
            for p in self.module.parameters():
                dist.broadcast(p.data, src=0, group=self.process_group)
            if hasattr(self.module, "invalidate_shadows"):
                self.module.invalidate_shadows()
            self._needs_broadcast = False
        return self.module(*args, **kwargs)

    @contextlib.contextmanager
    def no_sync(self):
        old, self._require_sync = self._require_sync, False
        try:
            yield
        finally:
            self._require_sync = old

    # -- backward-time hooks -------------------------------------------------------------------------
    def _on_ready(self, p):
        if not self._require_sync or self.world_size == 1:
            return
        if not self._callback_queued:
            Variable._execution_engine.queue_callback(self._finalize_backward)
            self._callback_queued = True
        b = p._tae_bucket
        if id(p) in b.seen:
            # a parameter used twice in one forward reports twice; its bucket must not have left yet
            if b.work is not None:
                raise RuntimeError("tae_b200.DistributedDataParallel: a gradient arrived after its bucket had been "
                                   "all-reduced (parameter shared between modules?) — replicas would diverge")
            return
        b.seen.add(id(p))
        b.pending -= 1
        if b.pending == 0 and self.overlap:
            self._launch(b)

    def _launch(self, b):
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            from .tae import wgrad_side_stream

            side = wgrad_side_stream(b.flat.device, busy_only=True)  # weight gradients of short-grid models: side stream
            if side is not None:
                self._comm_stream.wait_stream(side)
            with torch.cuda.stream(self._comm_stream):
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.process_group, async_op=True)
        else:  # CPU / gloo (tests): gloo has no AVG
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
        self._launched.append(b)

    def _finalize_backward(self):
        # buckets whose parameters did not all report (unused parameters) are reduced now — and, without overlap, all
        for b in self._buckets:
            if b.work is None and b.pending != len(b.params):
                self._launch(b)
        for b in self._launched:
            if self._comm_stream is not None:
                b.work.wait()  # stream-level wait: orders the CURRENT stream after the collective, no host sync
            else:
                b.work.wait()
                b.flat.div_(self.world_size)
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._launched.clear()
        self._callback_queued = False
        # re-arm here, not in forward(): a training forward through the unwrapped module, or several forwards before
        # one backward, must find the bookkeeping fresh (every backward that produced a gradient ends in this callback)
        for b in self._buckets:
            b.rearm()

    # -- nn.Module plumbing so that checkpoints hold the unwrapped names --------------------------------
    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.load_state_dict(*args, **kwargs)
