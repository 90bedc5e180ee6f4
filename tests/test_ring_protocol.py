"""Discrete-event model of the barrier protocol of `attn_fwd_tc_ring` (tae_b200/csrc/attention_sm100.cu).

The kernel has four kinds of agents per CTA — TMA producer, MMA issuer, and one softmax group per query tile — that hand
TMEM ring buffers, the two O accumulators and the two operand buffers to each other through mbarriers waited on by PHASE
PARITY.  A wrong parity expression or a missing wait does not show up as a compile error, only as a hang or as silently
wrong numbers on the GPU, so the protocol is restated here with the kernel's own index arithmetic and run under randomised
latencies.  Checked: no deadlock, every wait returns for the phase it was meant for (no parity aliasing), a ring buffer /
O accumulator / operand buffer is never overwritten while somebody still needs its contents.

This is a model of the host-visible logic, not of the arithmetic; it runs on the CPU.
"""
import heapq
import itertools
import random

import pytest


class MBar:
    """mbarrier with an arrival count, waited on by parity (mbarrier.try_wait.parity semantics)."""

    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier expects"
        if self.pending == 0:
            self.pending = self.count
            self.completed += 1

    def ready(self, parity):
        # the phase with this parity has completed <=> the phase in progress has the other parity
        return (self.completed & 1) != parity


class Sim:
    def __init__(self, items, seed):
        self.rng = random.Random(seed)
        self.items = items
        self.now = 0.0
        self.events = []  # (time, seq, callable)
        self.seq = itertools.count()
        B = lambda n, c: MBar(n, c)
        self.kq = [B(f"kq{i}", 1) for i in range(2)]
        self.v = [B(f"v{i}", 1) for i in range(2)]
        self.buffree = [B(f"buffree{i}", 2) for i in range(2)]
        self.s = [B(f"s{i}", 1) for i in range(3)]
        self.p = [B(f"p{i}", 1) for i in range(3)]      # 256 thread arrivals in the kernel, one group here
        self.pv = [B(f"pv{i}", 1) for i in range(3)]
        self.ofree = [B(f"ofree{i}", 1) for i in range(2)]
        # resource state, for the safety checks
        self.ring = [None] * 3       # ("S", g) scores of job g | ("P", g) probabilities | ("free", g) consumed by PV(g)
        self.o_owner = [None] * 2    # (item, "lo"/"hi"/"read")
        self.opbuf = [None] * 2      # item whose operands are resident
        self.opbuf_loading = [None] * 2
        self.mma_queue_free_at = 0.0
        self.log = []

    # --- event machinery ---------------------------------------------------------------------
    def at(self, dt, fn):
        heapq.heappush(self.events, (self.now + dt, next(self.seq), fn))

    def lat(self, lo, hi):
        return self.rng.uniform(lo, hi)

    def run(self, agents):
        waiting = list(agents)
        # agents are generators yielding ("wait", barrier, parity, expected_completed) or ("sleep", dt)
        blocked = {}
        ready = [(a, None) for a in waiting]
        done = set()
        while True:
            progressed = False
            for a, _ in ready:
                self._step(a, blocked, done)
                progressed = True
            ready = []
            # re-check blocked agents
            for a, (bar, parity, expect) in list(blocked.items()):
                if bar.ready(parity):
                    assert bar.completed == expect, (f"parity aliasing on {bar.name}: wait meant for completion #{expect} "
                                                     f"returned at #{bar.completed}")
                    del blocked[a]
                    ready.append((a, None))
            if ready:
                continue
            if not self.events:
                break
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            r = fn()
            if r is not None:  # a sleeping agent resumes
                ready.append((r, None))
        assert not blocked, "deadlock: " + ", ".join(f"{a.__name__ if hasattr(a, '__name__') else a} on {b.name}"
                                                     for a, (b, _, _) in blocked.items())
        assert len(done) == len(agents), "an agent did not finish"

    def _step(self, a, blocked, done):
        while True:
            try:
                req = next(a)
            except StopIteration:
                done.add(a)
                return
            if req[0] == "wait":
                _, bar, parity, expect = req
                if bar.ready(parity):
                    assert bar.completed == expect, (f"parity aliasing on {bar.name}: wait meant for completion #{expect} "
                                                     f"returned at #{bar.completed}")
                    continue
                blocked[a] = (bar, parity, expect)
                return
            if req[0] == "sleep":
                self.at(req[1], lambda a=a: a)
                return

    # --- asynchronous engines ------------------------------------------------------------------
    def tma_load(self, s, item, bar_kq, bar_v):
        assert self.opbuf[s] is None and self.opbuf_loading[s] is None, f"operand buffer {s} reloaded while item {self.opbuf[s]} is resident"
        self.opbuf_loading[s] = item

        def land_kq():
            bar_kq.arrive()

        def land_v():
            self.opbuf[s], self.opbuf_loading[s] = item, None
            bar_v.arrive()

        d = self.lat(300, 3000)
        self.at(d, land_kq)
        self.at(d + self.lat(0, 800), land_v)

    def mma(self, dur, on_start, on_done):
        # the tensor core executes the issued MMAs one after the other
        start = max(self.now, self.mma_queue_free_at)
        self.mma_queue_free_at = start + dur
        self.at(start - self.now, on_start)
        self.at(start + dur - self.now, on_done)


def producer(sim):
    for it in range(sim.items):
        s = it & 1
        if it >= 2:
            yield ("wait", sim.buffree[s], ((it >> 1) - 1) & 1, it >> 1)
            assert sim.opbuf[s] == "released", f"producer reloads buffer {s} before item {it - 2} released it"
            sim.opbuf[s] = None
        sim.tma_load(s, it, sim.kq[s], sim.v[s])
        yield ("sleep", sim.lat(5, 20))


def issuer(sim):
    G = 4 * sim.items

    def issue_s(g):
        it, j = g >> 2, g & 3
        t, hi, s, rb = j & 1, j >> 1, it & 1, g % 3
        yield ("wait", sim.kq[s], (it >> 1) & 1, (it >> 1) + 1)
        if g >= 3:
            yield ("wait", sim.pv[rb], ((g - 3) // 3) & 1, (g - 3) // 3 + 1)

        def start():
            st = sim.ring[rb]
            assert st is None or st == ("free", g - 3), f"S({g}) overwrites ring buffer {rb} holding {st}"
            assert sim.opbuf[s] == it or sim.opbuf_loading[s] == it, f"S({g}) reads operands of item {sim.opbuf[s]}"
            sim.ring[rb] = ("S*", g)

        def done():
            sim.ring[rb] = ("S", g)
            sim.s[rb].arrive()

        sim.mma(sim.lat(250, 400), start, done)
        yield ("sleep", sim.lat(10, 40))

    def issue_pv(g):
        it, j = g >> 2, g & 3
        t, hi, s, rb = j & 1, j >> 1, it & 1, g % 3
        yield ("wait", sim.v[s], (it >> 1) & 1, (it >> 1) + 1)
        if not hi and it >= 1:
            yield ("wait", sim.ofree[t], (it - 1) & 1, it)
        yield ("wait", sim.p[rb], (g // 3) & 1, g // 3 + 1)

        def start():
            assert sim.ring[rb] == ("P", g), f"PV({g}) reads ring buffer {rb} holding {sim.ring[rb]}"
            assert sim.opbuf[s] == it, f"PV({g}) reads V of item {sim.opbuf[s]}"
            if not hi:
                assert sim.o_owner[t] is None or sim.o_owner[t] == (it - 1, "read"), f"PV({g}) overwrites O_{t} = {sim.o_owner[t]}"
                sim.o_owner[t] = (it, "lo*")
            else:
                assert sim.o_owner[t] == (it, "lo"), f"PV({g}) accumulates onto O_{t} = {sim.o_owner[t]}"
                sim.o_owner[t] = (it, "hi*")

        def done():
            sim.ring[rb] = ("free", g)
            sim.o_owner[t] = (it, "hi" if hi else "lo")
            sim.pv[rb].arrive()

        sim.mma(sim.lat(350, 600), start, done)
        yield ("sleep", sim.lat(10, 40))

    if G > 0:
        yield from issue_s(0)
    if G > 1:
        yield from issue_s(1)
    for g in range(G):
        if g + 2 < G:
            yield from issue_s(g + 2)
        yield from issue_pv(g)


def group(sim, t, slow_path_prob):
    def softmax_job(g, hi):
        rb = g % 3
        yield ("wait", sim.s[rb], (g // 3) & 1, g // 3 + 1)
        assert sim.ring[rb] == ("S", g), f"group {t} reads ring buffer {rb} for job {g} but it holds {sim.ring[rb]}"
        yield ("sleep", sim.lat(200, 500))  # max pass + exchange
        if hi and sim.rng.random() < slow_path_prob:
            yield ("wait", sim.pv[(g - 2) % 3], ((g - 2) // 3) & 1, (g - 2) // 3 + 1)
            assert sim.o_owner[t] == (g >> 2, "lo"), f"rescale of O_{t} = {sim.o_owner[t]}"
            yield ("sleep", sim.lat(100, 300))
        yield ("sleep", sim.lat(800, 2200))  # exp pass
        assert sim.ring[rb] == ("S", g)
        sim.ring[rb] = ("P", g)
        sim.p[rb].arrive()

    if sim.items > 0:
        yield from softmax_job(t, 0)
    for it in range(sim.items):
        s = it & 1
        g_hi = 4 * it + 2 + t
        yield from softmax_job(g_hi, 1)
        if it + 1 < sim.items:
            yield from softmax_job(4 * (it + 1) + t, 0)
        # epilogue of item `it`
        yield ("sleep", sim.lat(50, 150))
        yield ("wait", sim.pv[g_hi % 3], (g_hi // 3) & 1, g_hi // 3 + 1)
        assert sim.o_owner[t] == (it, "hi"), f"epilogue of item {it} reads O_{t} = {sim.o_owner[t]}"
        sim.o_owner[t] = (it, "read")
        sim.ofree[t].arrive()
        yield ("sleep", sim.lat(200, 600))  # normalise, stage over the dead Q tile, TMA store, wait for its read
        assert sim.opbuf[s] == it
        sim.buffree[s].arrive()
        if sim.buffree[s].pending == sim.buffree[s].count:  # both tiles have arrived: the buffer is released
            sim.opbuf[s] = "released"


@pytest.mark.parametrize("items", [1, 2, 3, 4, 7])
def test_ring_attention_protocol(items):
    for seed in range(60):
        for slow in (0.0, 0.5, 1.0):
            sim = Sim(items, seed * 7 + int(slow * 2))
            sim.run([producer(sim), issuer(sim), group(sim, 0, slow), group(sim, 1, slow)])
            assert all(st is None or st[0] == "free" for st in sim.ring)
            assert all(o == (items - 1, "read") for o in sim.o_owner)
