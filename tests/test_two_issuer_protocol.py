"""CPU: discrete-event model of a TMA ring consumed by TWO MMA issuer warps (round 2: every tcgen05 kernel splits its MMA stream
over two issuers, genesis_b200/csrc/igemm_halo.cu, wgrad_tc.cu).

mbarriers follow the PTX semantics: try_wait.parity(P) succeeds once the phase of parity P has completed, i.e. it only means
"the phase I am waiting for" to a waiter that is within ONE phase of the barrier.  Three release policies are modelled with
random latencies and occasional long stalls of one issuer (what a co-running kernel does to a warp):

  shared       both issuers wait for every stage and both release it (barrier count 2)  -- conv_halo_kernel /
               conv_halo_persistent_kernel (fullB / emptyB, fullA / emptyA), wgrad_halo_kernel (full / empty).  Must hold.
  owner_skips  an issuer waits for and releases only the stages it owns (count 1) -- the first two-issuer version of the tile
               weight-gradient kernel: exact in every isolated test, deadlocked inside the training step.  Must be flagged.
  owner_rel    both wait, only the owner releases (count 1) -- the intermediate fix: the non-owner can still fall a phase
               behind.  Must be flagged.

The tile weight-gradient kernel went back to ONE issuer (it is L2-bound; the split bought nothing)."""
import heapq
import random

import pytest


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


class Ring(object):
    def __init__(self, policy, n_iters, stages, n_groups, seed, stall=0.0):
        self.policy, self.n, self.S, self.ng = policy, n_iters, stages, n_groups
        self.rng = random.Random(seed)
        self.stall = stall
        self.full = [Bar(1) for _ in range(stages)]
        self.empty = [Bar(2 if policy == 'shared' else 1) for _ in range(stages)]
        self.content = [None] * stages
        self.reading = [0] * stages
        self.errors, self.consumed = [], []
        self.t, self.events, self.seq = 0.0, [], 0

    def at(self, dt, fn):
        self.seq += 1
        heapq.heappush(self.events, (self.t + dt, self.seq, fn))

    def producer(self):
        for it in range(self.n):
            s = it % self.S
            yield (self.empty[s], ((it // self.S) & 1) ^ 1)

            def land(s=s, it=it):
                if self.reading[s]:
                    self.errors.append('stage %d overwritten by load %d while it is being read' % (s, it))
                self.content[s] = it
                self.full[s].arrive()
            self.at(self.rng.uniform(3, 40), land)

    def issuer(self, iw):
        for it in range(self.n):
            s, g = it % self.S, it % self.ng
            mine = (g & 1) == iw
            if self.policy == 'owner_skips' and not mine:
                continue
            if self.rng.random() < self.stall:
                yield self.rng.uniform(100, 400)              # a co-running kernel holds this warp back
            yield (self.full[s], (it // self.S) & 1)
            if mine:
                if self.content[s] != it:
                    self.errors.append('issuer %d expected load %d in stage %d, found %r' % (iw, it, s, self.content[s]))
                self.reading[s] += 1
                self.consumed.append(it)

                def retire(s=s):
                    self.reading[s] -= 1
                    self.empty[s].arrive()                    # tcgen05.commit: arrives when the MMAs have retired
                self.at(self.rng.uniform(2, 30), retire)
            elif self.policy == 'shared':
                self.empty[s].arrive()                        # plain arrive / commit without MMAs of its own
            yield self.rng.uniform(0.5, 3)

    def run(self):
        roles = [self.producer(), self.issuer(0), self.issuer(1)]
        waiting = {i: ('ready', None) for i in range(3)}
        live = set(range(3))
        steps = 0
        while live:
            steps += 1
            assert steps < 1_000_000, 'runaway simulation'
            progressed = False
            for i in sorted(live):
                kind, arg = waiting[i]
                if not (kind == 'ready' or (kind == 'bar' and arg[0].done(arg[1])) or (kind == 'sleep' and self.t >= arg)):
                    continue
                progressed = True
                try:
                    y = next(roles[i])
                except StopIteration:
                    live.discard(i)
                    continue
                waiting[i] = ('bar', y) if isinstance(y, tuple) else ('sleep', self.t + y)
            if progressed:
                continue
            times = [a for k, (kind, a) in waiting.items() if k in live and kind == 'sleep']
            if self.events:
                times.append(self.events[0][0])
            if not times:
                raise AssertionError('deadlock')
            self.t = max(self.t, min(times))
            while self.events and self.events[0][0] <= self.t:
                heapq.heappop(self.events)[2]()
        while self.events:
            heapq.heappop(self.events)[2]()


SHAPES = [(60, 2, 3), (60, 3, 3), (64, 3, 13), (40, 4, 5), (30, 2, 1), (50, 8, 9)]      # iterations, ring stages, M-groups


@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('seed', range(4))
def test_shared_release_holds(shape, seed):
    n, S, ng = shape
    for stall in (0.0, 0.1, 0.4):
        r = Ring('shared', n, S, ng, seed, stall)
        r.run()
        assert not r.errors, r.errors[:3]
        assert sorted(r.consumed) == list(range(n))


@pytest.mark.parametrize('policy', ['owner_skips', 'owner_rel'])
def test_one_sided_release_is_flagged(policy):
    """Both broken policies must show a stale read, an overwrite or a deadlock for some timing -- and they pass for others, which
    is why the isolated GPU tests did not see them."""
    bad = good = 0
    for shape in SHAPES:
        n, S, ng = shape
        for seed in range(6):
            for stall in (0.0, 0.1, 0.4):
                r = Ring(policy, n, S, ng, seed, stall)
                try:
                    r.run()
                    ok = not r.errors and sorted(r.consumed) == list(range(n))
                except AssertionError:
                    ok = False
                bad += not ok
                good += ok
    assert bad > 0, 'the model no longer flags %s' % policy
    print(policy, 'flagged in', bad, 'of', bad + good, 'runs')
