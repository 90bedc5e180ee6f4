"""Training-loop parity on the GPU:

  * data parallelism on hardware (train.py:102): an N-rank step over NCCL equals the 1-rank step on the concatenated
    batch, and the replicas stay BIT-identical over several optimizer steps (needs >= 2 GPUs, skipped otherwise);
  * HostBatchFeeder (train.py:134): ordering and buffer recycling with several batches in flight;
  * FusedAdamW: checkpoint round trip (util/misc.py:307-323, train.py:161-172) in torch.optim.AdamW's format, both
    directions, and a parameter whose gradient arrives through plain autograd (a torch-native head);
  * the whole step captured in a CUDA graph replays to the same losses as eager execution.
"""
import os
import socket
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(img_size=64, patch_size=8, in_chans=3, embed_dim=128, vocab_size=16, depth=2, num_heads=2,
          decoder_embed_dim=128, decoder_depth=2, decoder_num_heads=2, mlp_ratio=4.)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(seed=0, kw=KW):
    from tae_b200 import tae as T

    torch.manual_seed(seed)
    return T.TAE(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw).cuda()


# ----------------------------------------------------------------------------------------------------
# a19: DDP on hardware
# ----------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddp_worker(rank, world, port, q, kw, per_rank, bucket_mb):
    try:
        import torch.distributed as dist

        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from tae_b200 import engine, misc
        from tae_b200.ddp import DistributedDataParallel

        S = kw["img_size"]
        xs = [torch.randn(world * per_rank, 3, S, S, generator=torch.Generator().manual_seed(100 + s)).cuda() for s in range(3)]

        # 1-rank reference on the concatenated batch (every rank computes it: same seed, same data)
        ref = build(0, kw)
        ref_opt = engine.build_optimizer(ref, max_lr=1e-3, weight_decay=0.05)
        loss_ref, _ = ref(xs[0])
        loss_ref.backward()
        ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}

        # N ranks, different initial weights per rank: the constructor's broadcast must equalise them
        model = build(rank, kw)
        opt = engine.build_optimizer(model, max_lr=1e-3, weight_decay=0.05)
        ddp = DistributedDataParallel(model, optimizer=opt, bucket_mb=bucket_mb)
        shard = slice(rank * per_rank, (rank + 1) * per_rank)
        loss, _ = ddp(xs[0][shard].contiguous())
        loss.backward()
        torch.cuda.synchronize()
        assert len(ddp._buckets) >= 2, len(ddp._buckets)
        worst = max((rel(p.grad, ref_grads[n]), n) for n, p in model.named_parameters())
        # same bf16 operands, fp32 accumulation in a different order (rows split across ranks, then averaged)
        assert worst[0] < 2e-4, worst
        lsum = loss.detach().clone()
        dist.all_reduce(lsum)
        assert abs(float(lsum) / world - float(loss_ref)) < 1e-5 * abs(float(loss_ref))

        # three optimizer steps: replicas bit-identical, and tracking the 1-rank run
        scaler = misc.NativeScalerWithGradNormCount(compute_norm=False)
        opt.step()
        opt.zero_grad()
        ref_opt.step()
        ref_opt.zero_grad()
        for it in (1, 2):
            l, _ = ddp(xs[it][shard].contiguous())
            scaler(l, opt)
            opt.zero_grad()
            lr_, _ = ref(xs[it])
            scaler(lr_, ref_opt)
            ref_opt.zero_grad()
            lsum = l.detach().clone()
            dist.all_reduce(lsum)
            assert abs(float(lsum) / world - float(lr_)) < 2e-3 * abs(float(lr_)), (it, float(lsum) / world, float(lr_))
        torch.cuda.synchronize()
        for ar in opt.arenas:
            for t in (ar.p, ar.pb, ar.m, ar.v):  # fp32 masters, bf16 shadows, both moments
                gathered = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(gathered, t)
                assert all(torch.equal(gathered[0], g) for g in gathered[1:]), "replicas diverged"
        # gradient accumulation: no_sync() micro-step + synced micro-step == one step on both (train.py:137-148)
        with ddp.no_sync():
            l, _ = ddp(xs[0][shard].contiguous())
            (l / 2).backward()
        l, _ = ddp(xs[1][shard].contiguous())
        (l / 2).backward()
        la, _ = ref(xs[0])
        (la / 2).backward()
        lb, _ = ref(xs[1])
        (lb / 2).backward()
        torch.cuda.synchronize()
        worst = max((rel(p.grad, dict(ref.named_parameters())[n].grad), n) for n, p in model.named_parameters()
                    if float(p.grad.norm()) > 1e-6)
        assert worst[0] < 5e-3, worst  # the two runs' weights already differ by fp32 reduction noise through 3 Adam steps
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import sys
        import traceback

        traceback.print_exc(file=sys.stderr)  # pytest's assertion repr truncates the queued message
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("kw,per_rank,bucket_mb", [
    (KW, 4, 0.25),
    (dict(img_size=256, patch_size=16, in_chans=3, embed_dim=256, vocab_size=64, depth=1, num_heads=4,
          decoder_embed_dim=256, decoder_depth=1, decoder_num_heads=4, mlp_ratio=4.), 2, 1.0),
])
def test_ddp_nccl_matches_single_rank_and_replicas_stay_identical(kw, per_rank, bucket_mb):
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q, kw, per_rank, bucket_mb)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    try:
        results = [q.get(timeout=200) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():  # a rank stuck in a collective its peer never entered must not outlive the test
                p.kill()
    assert len(results) == world and all(msg == "ok" for _, msg in results), results


# ----------------------------------------------------------------------------------------------------
# f1: host feeder
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("depth", [2, 3])
def test_host_batch_feeder_order_and_recycling(depth):
    """Seven distinct host batches through a ring of `depth` device buffers: every batch arrives intact and in order even
    though the copy of batch i+1 is in flight while batch i is being consumed by a slow kernel, and a buffer is only
    overwritten after the step that read it was released."""
    from tae_b200.engine import HostBatchFeeder

    host = [torch.full((8, 3, 64, 64), float(i + 1)).pin_memory() for i in range(7)]
    for i, h in enumerate(host):
        h[0, 0, 0, :8] = torch.arange(8) + 10 * i
    feeder = HostBatchFeeder(host, depth=depth)
    assert feeder.bytes_per_batch == 8 * 3 * 64 * 64 * 4
    sums, firsts, held = [], [], []
    spin = torch.randn(2048, 2048, device="cuda")
    for i in range(16):
        x = feeder.next()
        held.append(x.data_ptr())
        # a slow consumer on the compute stream: the next copy must not land in this buffer before release()
        for _ in range(3):
            spin = spin @ spin * 1e-3
        sums.append(x.sum())
        firsts.append(x[0, 0, 0, :8].clone())
        feeder.release()
    torch.cuda.synchronize()
    for i in range(16):
        k = i % 7
        want = float(host[k].sum())
        assert abs(float(sums[i]) - want) < 1e-3 * want, (i, float(sums[i]), want)
        assert torch.equal(firsts[i].cpu(), torch.arange(8.) + 10 * k)
    assert len(set(held)) == depth  # a ring of exactly `depth` device buffers


# ----------------------------------------------------------------------------------------------------
# f4: optimizer checkpoint round trip; a16/a17 with a foreign parameter
# ----------------------------------------------------------------------------------------------------
def test_fused_adamw_state_dict_roundtrip_and_torch_adamw_interchange(tmp_path):
    from tae_b200 import engine, misc

    xs = [torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(s)).cuda() for s in range(5)]
    scaler = misc.NativeScalerWithGradNormCount(compute_norm=False)

    def run(model, opt, batches):
        out = []
        for x in batches:
            l, _ = model(x)
            scaler(l, opt)
            opt.zero_grad()
            out.append(float(l))
        return out

    # uninterrupted run: 5 steps
    m0 = build()
    o0 = engine.build_optimizer(m0, max_lr=1e-3, weight_decay=0.05)
    full = run(m0, o0, xs)

    # 3 steps, checkpoint exactly as train.py:161-172 does, resume in fresh objects through misc.load_model, 2 more steps
    m1 = build()
    o1 = engine.build_optimizer(m1, max_lr=1e-3, weight_decay=0.05)
    first = run(m1, o1, xs[:3])
    import argparse

    ckpt = str(tmp_path / "ckpt.pth")
    misc.save_on_master({"model": m1.state_dict(), "optimizer": o1.state_dict(), "args": argparse.Namespace(model="tiny"),
                         "iteration": 3, "scaler": scaler.state_dict()}, ckpt)
    m2 = build(seed=123)  # different init: everything must come from the checkpoint
    o2 = engine.build_optimizer(m2, max_lr=1e-3, weight_decay=0.05)
    misc.load_model(ckpt, m2, optimizer=o2, loss_scaler=scaler, optim_resume=True)
    # everything came back bit for bit: parameters, bf16 shadows, both moments, the step counters
    for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(p, q), n
    for a1, a2 in zip(o1.arenas, o2.arenas):
        assert torch.equal(a1.m, a2.m) and torch.equal(a1.v, a2.v)
    assert o2._steps == o1._steps == [3, 3]
    resumed = run(m2, o2, xs[3:])
    for a2 in o2.arenas:  # the bf16 weights the GEMMs read were refreshed from the loaded masters and kept current since
        assert torch.equal(a2.pb, a2.p.to(torch.bfloat16))
    # the continuation follows the uninterrupted run (split-K weight gradients use fp32 atomics, so two runs of the same
    # step differ in the last bits: compare within fp32 reduction noise, not bit for bit)
    close = lambda a, b: all(abs(x - y) <= 2e-6 * abs(y) for x, y in zip(a, b))
    assert close(first, full[:3]), (first, full[:3])
    assert close(resumed, full[3:]), (resumed, full[3:])
    for (n, p), (_, q) in zip(m0.named_parameters(), m2.named_parameters()):
        if not n.endswith("attn.qkv.bias"):  # its key third has a zero true gradient: Adam amplifies rounding noise there
            assert rel(q, p) < 1e-4, n

    # a torch.optim.AdamW checkpoint loads into FusedAdamW (and the other way round): same format, same param order
    m3, m4 = build(), build()
    t3 = torch.optim.AdamW(misc.add_weight_decay(m3, 0.05), lr=1e-3, betas=(0.9, 0.95))
    for x in xs[:3]:
        l, _ = m3(x)
        l.backward()
        t3.step()
        t3.zero_grad(set_to_none=True)
    sd = t3.state_dict()
    m4.load_state_dict(m3.state_dict())
    o4 = engine.build_optimizer(m4, max_lr=1e-3, weight_decay=0.05)
    o4.load_state_dict(sd)
    back = o4.state_dict()
    assert [g["params"] for g in back["param_groups"]] == [g["params"] for g in sd["param_groups"]]
    for k, st in sd["state"].items():
        assert torch.equal(back["state"][k]["exp_avg"], st["exp_avg"]) and torch.equal(back["state"][k]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(back["state"][k]["step"]) == float(st["step"]) == 3.0
    # one more step on each side from the same state: same update within fp32 rounding of the two AdamW kernels
    l4, _ = m4(xs[3])
    scaler(l4, o4)
    l3, _ = m3(xs[3])
    l3.backward()
    t3.step()
    assert abs(float(l3) - float(l4)) < 1e-6 * abs(float(l3))
    for (n, p), (_, q) in zip(m3.named_parameters(), m4.named_parameters()):
        assert rel(q, p) < 1e-5, n
    # and FusedAdamW's state loads into torch.optim.AdamW
    t5 = torch.optim.AdamW(misc.add_weight_decay(m3, 0.05), lr=1e-3, betas=(0.9, 0.95))
    t5.load_state_dict(o4.state_dict())
    assert float(t5.state_dict()["state"][0]["step"]) == 4.0


def test_fused_adamw_with_a_torch_native_parameter():
    """a parameter whose gradient arrives through autograd's AccumulateGrad (not tae_b200's backward) must
    train; a parameter used by BOTH paths must receive the sum; nothing may be added onto stale arena contents."""
    from tae_b200 import misc
    from tae_b200.optim import FusedAdamW

    class WithHead(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.tae = build()
            self.head = torch.nn.Linear(16, 3).cuda()          # torch-native: gradients come through AccumulateGrad

        def forward(self, x, shared=False):
            loss, pred, latent = self.tae(x, return_latent=True)
            aux = self.head(latent.float()).pow(2).mean()
            if shared:  # a torch-native op that also reads a parameter the CUDA backward writes directly
                aux = aux + self.tae.decoder_pred.bias.pow(2).sum()
            return loss + aux

    def make(fused):
        torch.manual_seed(1)
        m = WithHead()
        groups = misc.add_weight_decay(m, 0.05)
        opt = FusedAdamW(groups, lr=1e-3, betas=(0.9, 0.95)) if fused else torch.optim.AdamW(groups, lr=1e-3, betas=(0.9, 0.95))
        return m, opt

    xs = [torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(s)).cuda() for s in range(3)]
    ma, oa = make(True)
    mb, ob = make(False)
    for it, x in enumerate(xs):
        la = ma(x, shared=it >= 1)
        la.backward()
        lb = mb(x, shared=it >= 1)
        lb.backward()
        for n in ("head.weight", "head.bias", "tae.decoder_pred.bias", "tae.decoder_pred.weight"):
            ga, gb = dict(ma.named_parameters())[n].grad, dict(mb.named_parameters())[n].grad
            # step 0: same weights on both sides; later steps compare two optimizers' trajectories (bf16 forward noise)
            assert ga is not None and rel(ga, gb) < (2e-3 if it == 0 else 1e-2), (it, n, rel(ga, gb))
        assert abs(float(misc.get_grad_norm_(ma.parameters())) - float(misc.get_grad_norm_(mb.parameters()))) < 1e-2 * float(misc.get_grad_norm_(mb.parameters()))
        oa.step()
        ob.step()
        oa.zero_grad()
        ob.zero_grad(set_to_none=True)
        assert all(p.grad is None for p in ma.parameters())
        assert abs(float(la) - float(lb)) < 5e-3 * abs(float(lb)), it
    for n in ("head.weight", "head.bias", "tae.decoder_pred.bias"):
        assert rel(dict(ma.named_parameters())[n], dict(mb.named_parameters())[n]) < 5e-3, n
    assert rel(ma.head.weight, torch.nn.Linear(16, 3).cuda().weight) > 1e-3  # and it really moved


# ----------------------------------------------------------------------------------------------------
# f1: the whole step as one CUDA graph
# ----------------------------------------------------------------------------------------------------
def test_graphed_train_step_replays_eager_losses():
    from tae_b200 import engine

    xs = [torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(s)).cuda() for s in range(6)]
    m0 = build()
    o0 = engine.build_optimizer(m0, max_lr=1e-3, weight_decay=0.05)
    sc = engine.misc.NativeScalerWithGradNormCount(compute_norm=False)
    eager = [float(engine.train_step(m0, o0, sc, x, it, max_lr=1e-3, min_lr=1e-4, switch_it=4)) for it, x in enumerate(xs)]

    m1 = build()
    o1 = engine.build_optimizer(m1, max_lr=1e-3, weight_decay=0.05)
    step = engine.GraphedTrainStep(m1, o1, xs[0], max_lr=1e-3, min_lr=1e-4, switch_it=4, warmup_steps=0)
    graphed = [float(step(x, it)) for it, x in enumerate(xs)]
    assert step.captured and step.launches_per_step > 50
    # same kernels in the same order on the same data (lr switch at it=4 included); split-K atomics leave last-bit noise
    assert all(abs(a - b) <= 2e-6 * abs(b) for a, b in zip(graphed, eager)), (graphed, eager)
    for (n, p), (_, q) in zip(m0.named_parameters(), m1.named_parameters()):
        if not n.endswith("attn.qkv.bias"):
            assert rel(q, p) < 1e-4, n


def test_graphed_encoder_replays_eager_latents():
    from tae_b200 import engine

    model = build().eval()
    xs = [torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(20 + s)).cuda() for s in range(5)]
    enc = engine.GraphedEncoder(model, xs[0], warmup_steps=1)
    for x in xs:
        z = enc(x).clone()
        assert torch.equal(z, engine.encode_batch(model, x))
    assert enc.graph is not None
    # a ragged last batch falls back to the eager path
    xr = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(30)).cuda()
    assert torch.equal(enc(xr), engine.encode_batch(model, xr))
