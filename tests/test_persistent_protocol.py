"""CPU: discrete-event model of the producer / MMA / epilogue protocol of conv_halo_persistent_kernel (igemm_halo.cu).

The kernel's three roles are replayed as coroutines with the SAME loop structure, buffer indices and mbarrier parities as the
CUDA source (activation-window counter u -> buffer u & 1, phase (u >> 1) & 1; accumulator set it & 1, phase (it >> 1) & 1;
weight ring stage / phase counters; resident weights waited with parity 0).  mbarriers follow the PTX semantics
(try_wait.parity(P) succeeds once the phase of parity P has completed); TMA loads and tcgen05.commit complete asynchronously
and in order.  The model checks, for many (items, channel blocks, taps, ring depth, resident) shapes and random latencies:
no deadlock; every MMA reads the window / weight tile that was loaded FOR IT (never a stale or overwritten one); no window
or weight stage is overwritten before the MMAs that read it have retired; the epilogue drains a complete accumulator set
and the MMA warp never writes a set that has not been drained.  It cannot prove the CUDA code right -- it pins the protocol
the code was written against, which is where persistent pipelines usually break."""
import heapq
import random

import pytest

MAX_CHUNKS = 8


class Bar(object):
    def __init__(self, count=1):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def done(self, parity):
        return (self.phase & 1) != parity


class Sim(object):
    def __init__(self, n_items, cblocks, ntaps, stages, resident, nch, seed, nch_items=None, skip_unneeded=False):
        self.rng = random.Random(seed)
        # nch_items[i]: chunks item i NEEDS (bottom-edge tiles need fewer); skip_unneeded: the producer loads only those
        self.nch_items = nch_items or [nch] * n_items
        self.skip_unneeded = skip_unneeded
        self.n_items, self.cblocks, self.ntaps, self.stages, self.resident, self.nch = n_items, cblocks, ntaps, stages, resident, nch
        self.t = 0.0
        self.events = []          # (time, seq, fn)
        self.seq = 0
        self.fullA = [[Bar() for _ in range(MAX_CHUNKS)] for _ in range(2)]
        self.emptyA = [Bar(), Bar()]
        self.fullB = [Bar() for _ in range(stages)]
        self.emptyB = [Bar() for _ in range(stages)]
        self.accFull = [Bar(), Bar()]
        self.accEmpty = [Bar(4), Bar(4)]
        self.A = [[None] * MAX_CHUNKS for _ in range(2)]       # content tags of window chunks
        self.B = [None] * stages
        self.acc = [None, None]                                 # (item, n_mmas_done, complete)
        self.mma_queue_free_at = 0.0                            # in-order tensor pipe
        self.inflight_reads_A = [[0] * MAX_CHUNKS for _ in range(2)]    # per window chunk
        self.inflight_reads_B = [0] * stages
        self.errors = []
        self.drained = []

    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.t + dt, self.seq, fn))

    # ---- asynchronous engines
    def tma_A(self, buf, j, tag):
        def land():
            if self.inflight_reads_A[buf][j]:
                self.errors.append('window %d chunk %d overwritten while %d MMAs still read it' % (buf, j, self.inflight_reads_A[buf][j]))
            self.A[buf][j] = tag
            self.fullA[buf][j].arrive()
        self.at(self.rng.uniform(5, 60), land)

    def tma_B(self, s, tag):
        def land():
            if self.inflight_reads_B[s]:
                self.errors.append('weight stage %d overwritten while MMAs still read it' % s)
            self.B[s] = tag
            self.fullB[s].arrive()
        self.at(self.rng.uniform(3, 40), land)

    def mma(self, buf, need_chunks, a_tag, s, b_tag, acc, item, first):
        """issue one tap's MMAs; they execute in order on the tensor pipe"""
        for j in range(need_chunks):
            if self.A[buf][j] != a_tag:
                self.errors.append('MMA of %r read window chunk %d holding %r' % (a_tag, j, self.A[buf][j]))
        if self.B[s] != b_tag:
            self.errors.append('MMA expected weights %r, stage holds %r' % (b_tag, self.B[s]))
        if first:
            if self.acc[acc] is not None and not self.acc[acc][2]:
                self.errors.append('accumulator set %d overwritten before it was drained' % acc)
            self.acc[acc] = [item, 0, False]
        elif self.acc[acc] is None or self.acc[acc][0] != item:
            self.errors.append('accumulating item %d into a set owned by %r' % (item, self.acc[acc]))
        for j in range(need_chunks):
            self.inflight_reads_A[buf][j] += 1
        self.inflight_reads_B[s] += 1
        start = max(self.t, self.mma_queue_free_at)
        self.mma_queue_free_at = start + self.rng.uniform(1, 6)

        def retire():
            for j in range(need_chunks):
                self.inflight_reads_A[buf][j] -= 1
            self.inflight_reads_B[s] -= 1
            self.acc[acc][1] += 1
        self.at(self.mma_queue_free_at - self.t, retire)

    def commit(self, bar):
        """tcgen05.commit: arrives when everything issued before it has retired (in-order pipe)"""
        self.at(max(self.t, self.mma_queue_free_at) - self.t + 0.001, bar.arrive)

    # ---- the three roles, mirroring the CUDA loops
    def producer(self):
        bi, bph, u, b_loaded = 0, 0, 0, False
        for item in range(self.n_items):
            for cb in range(self.cblocks):
                buf = u & 1
                yield (self.emptyA[buf], ((u >> 1) & 1) ^ 1)
                for j in range(self.nch_items[item] if self.skip_unneeded else self.nch):
                    self.tma_A(buf, j, (item, cb))
                if self.resident:
                    if not b_loaded:
                        for c2 in range(self.cblocks):
                            for tap in range(self.ntaps):
                                self.tma_B(c2 * self.ntaps + tap, (c2, tap))
                    b_loaded = True
                else:
                    for tap in range(self.ntaps):
                        s = bi
                        yield (self.emptyB[s], bph ^ 1)
                        self.tma_B(s, (item, cb, tap))
                        bi += 1
                        if bi == self.stages:
                            bi, bph = 0, bph ^ 1
                u += 1

    def mma_warp(self):
        bi, bph, u = 0, 0, 0
        for it in range(self.n_items):
            acc = it & 1
            yield (self.accEmpty[acc], ((it >> 1) & 1) ^ 1)
            for cb in range(self.cblocks):
                buf, aph = u & 1, (u >> 1) & 1
                waited = 0
                for tap in range(self.ntaps):
                    if self.resident:
                        s, ph, btag = cb * self.ntaps + tap, 0, (cb, tap)
                    else:
                        s, ph, btag = bi, bph, (it, cb, tap)
                    yield (self.fullB[s], ph)
                    # tap 0 stops one chunk short of the window's end; later taps (and a lone tap) reach the last chunk
                    nn = self.nch_items[it]
                    need = nn - 1 if (tap > 0 or self.ntaps == 1) else max(0, nn - 2)
                    while waited <= need:
                        yield (self.fullA[buf][waited], aph)
                        waited += 1
                    self.mma(buf, waited, (it, cb), s, btag, acc, it, first=(cb == 0 and tap == 0))
                    if not self.resident:
                        self.commit(self.emptyB[s])
                        bi += 1
                        if bi == self.stages:
                            bi, bph = 0, bph ^ 1
                while waited < (self.nch_items[it] if self.skip_unneeded else self.nch):     # chunks no tap needed have landed too
                    yield (self.fullA[buf][waited], aph)
                    waited += 1
                self.commit(self.emptyA[buf])
                u += 1
            done = self.accFull[acc]

            def complete(acc=acc, done=done):
                self.acc[acc][2] = 'mma-complete'
                done.arrive()
            self.at(max(self.t, self.mma_queue_free_at) - self.t + 0.001, complete)

    def epilogue_warp(self, w):
        for it in range(self.n_items):
            acc = it & 1
            yield (self.accFull[acc], (it >> 1) & 1)
            st = self.acc[acc]
            if st is None or st[0] != it or st[1] != self.cblocks * self.ntaps:
                self.errors.append('epilogue of item %d found accumulator state %r' % (it, st))
            yield self.rng.uniform(5, 80)                       # drain time
            if w == 0:
                self.drained.append(it)
            self.arrive_acc_empty(acc)

    def arrive_acc_empty(self, acc):
        bar = self.accEmpty[acc]
        before = bar.phase
        bar.arrive()
        if bar.phase != before:
            self.acc[acc][2] = True                             # all four warps done: the set may be overwritten

    def run(self):
        roles = [self.producer(), self.mma_warp()] + [self.epilogue_warp(w) for w in range(4)]
        waiting = {}                                            # role index -> (bar, parity) or wake time
        live = set(range(len(roles)))
        for i in list(live):
            waiting[i] = ('ready', None)
        steps = 0
        while live:
            steps += 1
            assert steps < 2_000_000, 'runaway simulation'
            progressed = False
            for i in sorted(live):
                kind, arg = waiting[i]
                ok = kind == 'ready' or (kind == 'bar' and arg[0].done(arg[1])) or (kind == 'sleep' and self.t >= arg)
                if not ok:
                    continue
                progressed = True
                try:
                    y = next(roles[i])
                except StopIteration:
                    live.discard(i)
                    continue
                if isinstance(y, tuple):
                    waiting[i] = ('bar', y)
                else:
                    waiting[i] = ('sleep', self.t + y)
            if progressed:
                continue
            # nobody can run: advance time to the next asynchronous completion or wake-up
            wake = [arg for k, (kind, arg) in waiting.items() if k in live and kind == 'sleep']
            nxt = min([self.events[0][0]] if self.events else [] + wake) if (self.events or wake) else None
            if self.events and wake:
                nxt = min(self.events[0][0], min(wake))
            if nxt is None:
                stuck = {i: waiting[i] for i in live}
                raise AssertionError('deadlock: %r' % {i: (k, (id(a[0]) % 1000, a[1]) if k == 'bar' else a) for i, (k, a) in stuck.items()})
            self.t = max(self.t, nxt)
            while self.events and self.events[0][0] <= self.t:
                _, _, fn = heapq.heappop(self.events)
                fn()
        while self.events:                                      # drain the tail
            tt, _, fn = heapq.heappop(self.events)
            self.t = max(self.t, tt)
            fn()


CONFIGS = [  # n_items, cblocks, ntaps, ring stages, resident, chunks per window
    (1, 1, 9, 9, True, 3),
    (7, 1, 9, 9, True, 3),          # 3x3 Cout 32: resident weights
    (22, 1, 9, 9, True, 4),
    (6, 2, 9, 18, True, 2),         # two channel blocks, resident
    (5, 1, 25, 4, False, 3),        # 5x5 Cout 64: ring of 4
    (9, 2, 25, 8, False, 5),        # 5x5 Cout 32, two channel blocks, ring of 8
    (4, 4, 9, 2, False, 1),         # UNet up block: four channel blocks, shallow ring
    (2, 1, 4, 2, False, 2),         # stride-2 conv-transpose class with 4 taps
    (3, 3, 1, 2, False, 1),         # 1x1 conv
]


@pytest.mark.parametrize('cfg', CONFIGS)
@pytest.mark.parametrize('seed', [0, 1, 2])
def test_persistent_protocol(cfg, seed):
    n_items, cblocks, ntaps, stages, resident, nch = cfg
    sim = Sim(n_items, cblocks, ntaps, stages, resident, nch, seed)
    sim.run()
    assert not sim.errors, sim.errors[:5]
    assert sim.drained == list(range(n_items))


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_ragged_items_with_all_chunks_loaded(seed):
    """Bottom-edge tiles need fewer window chunks; the kernel loads all of them anyway (zero-filled by TMA)."""
    nch_items = [4, 4, 2, 4, 3, 4, 4, 1, 4, 4]
    sim = Sim(len(nch_items), 1, 9, 9, True, 4, seed, nch_items=nch_items)
    sim.run()
    assert not sim.errors, sim.errors[:5]
    assert sim.drained == list(range(len(nch_items)))


def test_skipping_chunks_breaks_the_parity_scheme():
    """Why: if the producer skipped the chunks an item does not need (as the one-item kernel does), their barriers would fall
    one phase behind the (u >> 1) & 1 parity of the next full item, whose wait then passes on the stale phase and the MMA
    reads the previous window's chunk (or the protocol deadlocks).  The model must flag that policy."""
    nch_items = [4, 4, 2, 4, 4, 4, 4, 4]
    bad = 0
    for seed in range(4):
        sim = Sim(len(nch_items), 1, 9, 9, True, 4, seed, nch_items=nch_items, skip_unneeded=True)
        try:
            sim.run()
            bad += bool(sim.errors)
        except AssertionError:
            bad += 1
    assert bad == 4
