"""The reference's step loops restated around the B200 path (no data loading: callers hand in tensors).

  train_step        train.py:122-150   lr schedule -> forward -> backward (+ bucketed all-reduce) -> AdamW -> zero_grad
  evaluate_batch    train.py:203-223 / evaluate.py:84-102   forward under no_grad, loss only
  encode_batch      encode.py:80-88    forward_encoder under no_grad
  GraphedTrainStep  the same iteration captured once as a CUDA graph and replayed
  GraphedEncoder    encode_batch as a replayed CUDA graph
  HostBatchFeeder   train.py:134 / encode.py:82   pinned-host -> device copies, double-buffered on a copy stream
                    (the synchronous 201 MB H2D per step is the adjacent host overhead)
  shard_for_rank    batch sharding for multi-GPU encode / evaluate (no communication; rank r takes slice r)
  LatentWriter      encode.py:87,93-100   asynchronous device->host ring + background writer for the extracted latents
                    (the reference .cpu()-s every batch synchronously and concatenates at the end)
"""
from __future__ import annotations

import math
import queue
import sys
import threading

import torch

from . import misc
from . import tae as tae_models
from .optim import FusedAdamW


def build_model(name: str, device="cuda"):
    """tae.__dict__[args.model]() (train.py:94) + .to(device)."""
    model = tae_models.__dict__[name]()
    return model.to(device)


def build_optimizer(model, max_lr=1e-4, weight_decay=0.05, track_grad_norm=False):
    """train.py:108-109: two param groups from add_weight_decay, AdamW betas (0.9, 0.95)."""
    groups = misc.add_weight_decay(model, weight_decay, bias_wd=False)
    return FusedAdamW(groups, lr=max_lr, betas=(0.9, 0.95), track_grad_norm=track_grad_norm)


def train_step(model, optimizer, loss_scaler, samples, it: int, *, max_lr=1e-4, min_lr=1e-5, switch_it=450_000,
               accum_iter=1, check_finite=False):
    """One iteration of the hot loop (train.py:122-150).  Returns the (device) loss tensor of this micro-step.

    The reference reads `loss.item()` and calls `torch.cuda.synchronize()` every step (train.py:139,150); here both are
    opt-in (`check_finite`) so the host can run ahead of the device."""
    if it % accum_iter == 0:
        misc.adjust_learning_rate(optimizer, max_lr, min_lr, it, switch_it)
    loss, _ = model(samples)
    if check_finite:
        v = loss.item()
        if not math.isfinite(v):
            print("Loss is {}, stopping training".format(v))
            sys.exit(1)
    step_loss = loss / accum_iter if accum_iter != 1 else loss
    update = (it + 1) % accum_iter == 0
    loss_scaler(step_loss, optimizer, parameters=None, update_grad=update)
    if update:
        optimizer.zero_grad()
    return loss.detach()


class GraphedTrainStep:
    """train_step() captured ONCE as a CUDA graph (forward, loss, hand-written backward, fused AdamW, zero_grad) and
    replayed per iteration: ~1200 kernel launches (patch16) leave the host as one `cudaGraphLaunch`.

    What makes the step capturable: every kernel of the path is launched by the C ABI on the current stream with no
    allocation, synchronisation or host read-back; activations come from the graph's private memory pool; the optimizer's
    scalars (lr from adjust_learning_rate, bias corrections) live in device memory (`FusedAdamW.make_capturable`).  The
    first `warmup_steps` calls run eagerly (they are real training steps), the next one captures.  Single process: under
    tae_b200.ddp use the eager step.  (Capturing the bucketed NCCL all-reduces with the step works — measured +1 % at
    2 GPUs — but a replay interleaved with eager collectives hung one of the round-2 test runs, so it is not offered.)

        step = GraphedTrainStep(model, optimizer, example_batch, max_lr=1e-4, min_lr=1e-5, switch_it=450_000)
        for it, samples in enumerate(loader): loss = step(samples, it)      # device tensor, no host sync
    """

    def __init__(self, model, optimizer, example, *, max_lr=1e-4, min_lr=1e-5, switch_it=450_000, warmup_steps=3):
        if misc.get_world_size() > 1:
            raise RuntimeError("GraphedTrainStep is single-process; use train_step under tae_b200.ddp")
        self.model, self.optimizer = model, optimizer
        self.sched = (max_lr, min_lr, switch_it)
        self.warmup_steps = warmup_steps
        self.static_in = torch.empty_like(example)
        self.static_loss = None
        self.graph = None
        self.calls = 0
        self.launches_per_step = 0
        self._scaler = misc.NativeScalerWithGradNormCount(compute_norm=False)
        # warm-up and capture run on ONE dedicated stream: autograd replays a node's backward on the stream of its forward
        # and synchronises leaf streams at the end of backward, so a warm-up on another stream than the capturing one
        # would make the capture depend on uncaptured work
        self.stream = torch.cuda.Stream(device=example.device)
        optimizer.make_capturable()

    @property
    def captured(self) -> bool:
        return self.graph is not None

    def _capture(self):
        from . import ops

        model, opt = self.model, self.optimizer
        torch.cuda.synchronize()
        torch.cuda.empty_cache()  # the eager warm-up's cached activations and the graph's pool must not both stay resident
        n0 = ops.launch_count()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=self.stream):
                loss, _ = model(self.static_in)
                self._scaler(loss, opt, parameters=None, update_grad=True)
                opt.zero_grad()
                self.static_loss = loss.detach()
        except RuntimeError as e:
            if "uncaptured work" in str(e) or "capture" in str(e).lower():
                raise RuntimeError(
                    "CUDA-graph capture of the training step failed.  A common cause: an un-detached loss / output of an "
                    "EARLIER eager step of this model is still referenced — its autograd graph keeps the parameters' "
                    "AccumulateGrad nodes bound to the stream of that step, and autograd then synchronises the capture "
                    "with it.  Keep only `.detach()`-ed results of eager steps (engine.train_step returns them).") from e
            raise
        self.launches_per_step = ops.launch_count() - n0
        self.graph = g

    def __call__(self, samples, it: int):
        misc.adjust_learning_rate(self.optimizer, self.sched[0], self.sched[1], it, self.sched[2])
        self.calls += 1
        if self.calls <= self.warmup_steps:
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                loss, _ = self.model(samples)
                self._scaler(loss, self.optimizer, parameters=None, update_grad=True)
                self.optimizer.zero_grad()
                loss = loss.detach()
            cur.wait_stream(self.stream)
            samples.record_stream(self.stream)
            loss.record_stream(cur)
            return loss
        if self.graph is None:
            self._capture()
        self.static_in.copy_(samples, non_blocking=True)
        self.optimizer.prepare_step()
        self.graph.replay()
        return self.static_loss.clone()


class GraphedEncoder:
    """encode_batch() captured once as a CUDA graph and replayed (encode.py:80-88 is ~150 launches per batch of a few
    hundred microseconds each; there is no communication, so this also holds per rank under batch sharding).

        enc = GraphedEncoder(model, example_batch)
        for samples, targets in loader: writer.put(enc(samples.to(device, non_blocking=True)), targets)

    The returned latent is the graph's static output buffer: consume it (LatentWriter.put enqueues its copy on a side
    stream ordered after the replay) or clone it before the next call."""

    def __init__(self, model, example, warmup_steps: int = 2):
        self.model = model
        self.static_in = torch.empty_like(example)
        self.static_out = None
        self.graph = None
        self.calls = 0
        self.warmup_steps = warmup_steps

    @torch.no_grad()
    def __call__(self, samples):
        self.calls += 1
        if self.calls <= self.warmup_steps or samples.shape != self.static_in.shape:
            return self.model.forward_encoder(samples)  # warm-up, or a ragged last batch: eager
        if self.graph is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.static_out = self.model.forward_encoder(self.static_in)
            self.graph = g
        self.static_in.copy_(samples, non_blocking=True)
        self.graph.replay()
        return self.static_out


@torch.no_grad()
def evaluate_batch(model, samples):
    loss, _ = model(samples)
    return loss


@torch.no_grad()
def encode_batch(model, samples):
    return model.forward_encoder(samples)


def shard_for_rank(n_items: int, rank: int, world: int):
    """Contiguous, order-preserving shard [lo, hi) of n_items for `rank` of `world` (encode / evaluate sharding)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class HostBatchFeeder:
    """Streams host batches to the device on a dedicated copy stream, two batches in flight.

    `host_batches`: list of pinned fp32 host tensors that is cycled.  `next()` returns a device tensor whose copy
    has been ordered before the current stream; the buffer is recycled two calls later."""

    def __init__(self, host_batches, device="cuda", depth: int = 2):
        self.host = host_batches
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.bufs = [torch.empty_like(host_batches[0], device=self.device) for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.freed = [torch.cuda.Event() for _ in range(depth)]
        self.i = 0
        self.bytes_per_batch = host_batches[0].numel() * host_batches[0].element_size()
        for k in range(depth):
            self.freed[k].record(torch.cuda.current_stream(self.device))
        self._issue(0)

    def _issue(self, idx):
        slot = idx % self.depth
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.freed[slot])
            self.bufs[slot].copy_(self.host[idx % len(self.host)], non_blocking=True)
            self.ready[slot].record(self.stream)

    def next(self):
        slot = self.i % self.depth
        cur = torch.cuda.current_stream(self.device)
        self._issue(self.i + 1)  # prefetch the following batch while this one is consumed
        cur.wait_event(self.ready[slot])
        return self.bufs[slot]

    def release(self):
        """Call after the step that consumed the batch has been enqueued."""
        slot = self.i % self.depth
        self.freed[slot].record(torch.cuda.current_stream(self.device))
        self.i += 1


class LatentWriter:
    """Streams encoded latents off the device without stalling the encoder (encode.py:87,93-100).

    The reference does `latents.append(latent.cpu())` per batch (a synchronous D2H copy on the compute stream) and one
    `torch.save({"latents", "targets"})` at the end.  Here `put()` enqueues the copy on a dedicated stream into a ring
    of pinned host buffers and returns immediately; a background thread waits for each copy's event, moves the rows
    into pageable storage and frees the ring slot.  `close()` writes the SAME file layout as the reference
    (`{"latents": [n, N, V], "targets": [n]}`, encode.py:97-100), so downstream loaders are unaffected; with
    `shard_rows` the rows are written as `path.partNNNNN` files of that many rows instead (same dict per part)."""

    def __init__(self, path: str, device="cuda", depth: int = 3, shard_rows: int | None = None):
        self.path, self.device, self.depth, self.shard_rows = path, torch.device(device), depth, shard_rows
        self.stream = torch.cuda.Stream(device=self.device)
        self.pinned = [None] * depth
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.free = [threading.Semaphore(1) for _ in range(depth)]
        self.q: queue.Queue = queue.Queue()
        self.lat, self.tgt, self.parts, self.rows_pending = [], [], [], 0
        self.i = 0
        self.bytes_copied = 0
        self.err = None
        self.worker = threading.Thread(target=self._drain, daemon=True)
        self.worker.start()

    def put(self, latent: torch.Tensor, targets: torch.Tensor | None = None):
        """latent: device tensor [b, N, V] produced on the current stream; targets: host or device tensor [b] or None."""
        if self.err is not None:
            raise self.err
        slot = self.i % self.depth
        self.i += 1
        self.free[slot].acquire()  # blocks only when the host writer is `depth` batches behind
        buf = self.pinned[slot]
        if buf is None or buf.shape[1:] != latent.shape[1:] or buf.shape[0] < latent.shape[0] or buf.dtype != latent.dtype:
            buf = torch.empty(latent.shape, dtype=latent.dtype, pin_memory=True)
            self.pinned[slot] = buf
        produced = torch.cuda.Event()
        produced.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(produced)
            buf[:latent.shape[0]].copy_(latent, non_blocking=True)
            self.copied[slot].record(self.stream)
        latent.record_stream(self.stream)
        self.bytes_copied += latent.numel() * latent.element_size()
        self.q.put((slot, latent.shape[0], None if targets is None else targets.detach().cpu()))

    def _flush_part(self, final: bool):
        if not self.lat:
            return
        out = {"latents": torch.cat(self.lat), "targets": torch.cat(self.tgt) if self.tgt else torch.empty(0, dtype=torch.long)}
        name = self.path if (final and not self.parts and self.shard_rows is None) else f"{self.path}.part{len(self.parts):05d}"
        torch.save(out, name)
        self.parts.append(name)
        self.lat, self.tgt, self.rows_pending = [], [], 0

    def _drain(self):
        try:
            while True:
                item = self.q.get()
                if item is None:
                    break
                slot, rows, tg = item
                self.copied[slot].synchronize()
                self.lat.append(self.pinned[slot][:rows].clone())
                self.free[slot].release()
                if tg is not None:
                    self.tgt.append(tg)
                self.rows_pending += rows
                if self.shard_rows is not None and self.rows_pending >= self.shard_rows:
                    self._flush_part(False)
            self._flush_part(True)
        except Exception as e:  # surfaced by the next put()/close()
            self.err = e
            for s in self.free:
                s.release()

    def close(self):
        """Waits for every copy and write; returns the list of files written."""
        self.q.put(None)
        self.worker.join()
        if self.err is not None:
            raise self.err
        return list(self.parts)
