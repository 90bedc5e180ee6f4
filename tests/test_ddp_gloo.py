"""CPU test of the data-parallel bucket logic with the gloo backend, world_size 2 (no GPU needed).

The gradient arena, bucket planning, ready-callbacks, end-of-backward finalisation and the rank-0 parameter broadcast
of tae_b200.ddp are device-agnostic; here two processes run a fake "backward" that writes per-rank gradients into the
arena exactly the way the hand-written CUDA backward does (`p._tae_grad`, `p._tae_ready(p)`) and check that every
rank ends up with the rank-average, bucket by bucket."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _WriteGrads(torch.autograd.Function):
    """Stands in for the CUDA backward: writes gradients straight into the arena and reports readiness."""

    @staticmethod
    def forward(ctx, x, params, rank):
        ctx.params, ctx.rank = params, rank
        return x.sum()

    @staticmethod
    def backward(ctx, g):
        for i, p in enumerate(reversed(ctx.params)):  # gradient-ready order = reverse parameter order
            p._tae_grad.fill_(float(ctx.rank + 1) * (i + 1))
            p._tae_dirty = 1
            p._tae_ready(p)
        return torch.ones(1) * g, None, None


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from tae_b200.ddp import DistributedDataParallel
        from tae_b200.optim import _Arena

        torch.manual_seed(rank)  # different initial params per rank: the broadcast must equalise them
        lin = torch.nn.Sequential(torch.nn.Linear(40, 30), torch.nn.Linear(30, 7), torch.nn.LayerNorm(7))
        params = list(lin.parameters())
        ar = _Arena(params, torch.device("cpu"))
        with torch.no_grad():
            for p in params:
                ar.view(ar.p, p).copy_(p)
                p.data = ar.view(ar.p, p)
                p._tae_grad = ar.view(ar.g, p)
                p.grad = p._tae_grad
                p._tae_dirty = 0

        class Opt:
            arenas = [ar]

        ddp = DistributedDataParallel(lin, optimizer=Opt(), bucket_mb=0.0005, tail_mb=0.0003)  # ~128-element buckets -> several buckets
        assert len(ddp._buckets) >= 2
        # parameters broadcast from rank 0
        ref = [torch.empty_like(ar.p) for _ in range(world)]
        dist.all_gather(ref, ar.p)
        assert torch.equal(ref[0], ref[1])

        x = torch.ones(1, requires_grad=True)
        for it in range(3):  # several steps, never through ddp.forward(): the end-of-backward callback re-arms
            with torch.enable_grad():
                loss = _WriteGrads.apply(x, params, rank)
                loss.backward()
            assert all(b.work is None and b.pending == len(b.params) and not b.seen for b in ddp._buckets)
            mean_scale = sum(r + 1 for r in range(world)) / world
            for i, p in enumerate(reversed(params)):
                want = mean_scale * (i + 1)
                assert torch.allclose(p.grad, torch.full_like(p.grad, want)), (rank, i, p.grad.flatten()[:3], want)
        # overlap=False: every bucket leaves at the end of backward; same averages
        ddp.overlap = False
        with torch.enable_grad():
            _WriteGrads.apply(x, params, rank).backward()
        for i, p in enumerate(reversed(params)):
            assert torch.allclose(p.grad, torch.full_like(p.grad, mean_scale * (i + 1))), (rank, i)
        assert all(b.work is None and b.pending == len(b.params) for b in ddp._buckets)
        ddp.overlap = True
        # no_sync: gradients stay local
        with ddp.no_sync():
            _WriteGrads.apply(x, params, rank).backward()
        assert torch.allclose(params[-1].grad, torch.full_like(params[-1].grad, float(rank + 1)))
        # a parameter that reports twice in one backward: harmless while its bucket has not left, an error afterwards
        b0 = next(b for b in ddp._buckets if len(b.params) >= 2)
        first = b0.params[0]
        ddp._callback_queued = True  # outside a backward pass: finalised by hand below
        ddp._on_ready(first)
        ddp._on_ready(first)
        assert b0.pending == len(b0.params) - 1
        for p in b0.params[1:]:
            ddp._on_ready(p)
        assert b0.work is not None
        try:
            ddp._on_ready(first)
            raise AssertionError("late gradient was not detected")
        except RuntimeError:
            pass
        ddp._finalize_backward()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.timeout(180)
def test_bucketed_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(msg == "ok" for _, msg in results), results


def test_plan_buckets_covers_arena_and_cuts_a_short_tail():
    """Bucket planning: contiguous cover of the arena in gradient-ready order; with tail_elems the gradients that become
    ready last form a short bucket of their own (the one all-reduce nothing can overlap)."""
    from tae_b200.ddp import plan_buckets

    class P:
        def __init__(self, n):
            self.n = n

        def numel(self):
            return self.n

    ps = [P(n) for n in (1000, 64, 5000, 300, 10, 2000, 7, 90, 33)]
    off, o = {}, 0
    for p in ps:
        off[id(p)] = o
        o += (p.n + 63) // 64 * 64
    for tail in (0, 200, 2500, 10 ** 9):
        b = plan_buckets(ps, off, lambda p: p.numel(), 4000, tail)
        assert b[0][0] == 0 and b[-1][1] is None and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert [p for _, _, pl in b for p in pl] == ps
        if 0 < tail < 10 ** 9:
            start = b[-1][0]
            assert o - start <= tail                                  # the tail bucket respects its cap ...
            assert o - off[id(ps[ps.index(b[-1][2][0]) - 1])] > tail  # ... and is maximal
    assert [len(pl) for _, _, pl in plan_buckets(ps, off, lambda p: p.numel(), 4000, 200)] == [3, 4, 2]
