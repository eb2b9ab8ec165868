"""CPU model of the synchronisation protocol of the opt-in warp-specialised GEMM (csrc/tc_gemm2.cuh).

The kernel was written without GPU time to run it, and a wrong mbarrier parity there means a hung GPU box.  This test
transcribes the control flow of the three roles — producers, MMA issuer, epilogue — with the kernel's own ring indices
and phase parities into Python generators, runs them under many random interleavings against a model of the mbarrier
semantics (arrival counts, phase completion, `try_wait.parity`), and checks
  * no deadlock (every role finishes every tile),
  * no hazard: a smem stage is never refilled before the MMA that read it has retired, an accumulator buffer is never
    overwritten before the epilogue has drained it, a chunk-list buffer is never rewritten before the MMA warp has read
    its count, and the MMA always consumes the chunk the producers put there.
It models the protocol only (who waits for what), not the data path."""
import random

import pytest


class MBar:
    """mbarrier: `count` arrivals complete a phase; wait(parity) passes once the phase with that parity has completed."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0, 'more arrivals than the barrier expects in one phase'
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def passed(self, parity):
        return self.phase != parity


class Model:
    def __init__(self, sa, sb, tiles, chunks_of_tile, gather, rng):
        self.SA, self.SB, self.tiles, self.chunks, self.gather, self.rng = sa, sb, tiles, chunks_of_tile, gather, rng
        self.a_full = [MBar(1) for _ in range(sa)]       # 256 producer arrivals modelled as one
        self.a_empty = [MBar(1) for _ in range(sa)]
        self.b_full = [MBar(1) for _ in range(sb)]
        self.b_empty = [MBar(1) for _ in range(sb)]
        self.acc_full = [MBar(1), MBar(1)]
        self.acc_empty = [MBar(1), MBar(1)]              # 128 epilogue arrivals modelled as one
        self.list_full = [MBar(1), MBar(1)]
        self.list_empty = [MBar(1), MBar(1)]
        # data-path shadows
        self.a_stage = [None] * sa                       # (tile, chunk) currently in the stage, None = free
        self.b_stage = [None] * sb
        self.list_buf = [None, None]                     # (tile, count) or None once read
        self.acc = [None, None]                          # tile whose sum sits in the buffer, None = drained
        self.inflight = []                               # MMAs issued but not retired: list of callables (commits)
        self.done_tiles = []

    # tcgen05.commit: the arrival happens when all previously issued MMAs have retired -> model retirement as a
    # separate asynchronous step the scheduler may delay arbitrarily (but in order)
    def retire_one(self):
        if self.inflight:
            self.inflight.pop(0)()

    def producer(self):
        ca = 0
        for it, tile in enumerate(self.tiles):
            buf, ubuf = it & 1, it >> 1
            n_act = self.chunks[tile]
            if self.gather:
                if ubuf > 0:
                    while not self.list_empty[buf].passed((ubuf - 1) & 1):
                        yield
                assert self.list_buf[buf] is None, 'chunk-list buffer rewritten before the MMA warp read its count'
                self.list_buf[buf] = (tile, n_act)
                self.list_full[buf].arrive()
            for i in range(n_act):
                sa, sb = ca % self.SA, ca % self.SB
                ua, ub = ca // self.SA, ca // self.SB
                if ua > 0:
                    while not self.a_empty[sa].passed((ua - 1) & 1):
                        yield
                # thread 0 also feeds the weight ring
                if ub > 0:
                    while not self.b_empty[sb].passed((ub - 1) & 1):
                        yield
                assert self.b_stage[sb] is None, 'weight stage refilled before its MMA retired'
                self.b_stage[sb] = (tile, i)
                self.b_full[sb].arrive()                 # expect_tx arrive + TMA completion
                assert self.a_stage[sa] is None, 'A stage refilled before its MMA retired'
                self.a_stage[sa] = (tile, i)
                if self.rng.random() < 0.5:
                    yield                                # cp.async lands later
                self.a_full[sa].arrive()
                ca += 1
                yield

    def mma(self):
        ca = 0
        for it, tile in enumerate(self.tiles):
            buf, ubuf = it & 1, it >> 1
            n_act = self.chunks[tile]
            if self.gather:
                while not self.list_full[buf].passed(ubuf & 1):
                    yield
                assert self.list_buf[buf] == (tile, n_act), 'MMA warp read a stale chunk count'
                self.list_buf[buf] = None
                self.list_empty[buf].arrive()
            if ubuf > 0:
                while not self.acc_empty[buf].passed((ubuf - 1) & 1):
                    yield
            assert self.acc[buf] is None, 'accumulator overwritten before the epilogue drained it'
            for q in range(n_act):
                sa, sb = ca % self.SA, ca % self.SB
                while not self.a_full[sa].passed((ca // self.SA) & 1):
                    yield
                while not self.b_full[sb].passed((ca // self.SB) & 1):
                    yield
                assert self.a_stage[sa] == (tile, q) and self.b_stage[sb] == (tile, q), 'MMA consumed the wrong chunk'

                def commit(sa=sa, sb=sb, last=(q == n_act - 1), buf=buf, tile=tile):
                    self.a_stage[sa] = None
                    self.b_stage[sb] = None
                    self.a_empty[sa].arrive()
                    self.b_empty[sb].arrive()
                    if last:
                        self.acc[buf] = tile
                        self.acc_full[buf].arrive()
                self.inflight.append(commit)
                ca += 1
                yield

    def epilogue(self):
        for it, tile in enumerate(self.tiles):
            buf, ubuf = it & 1, it >> 1
            while not self.acc_full[buf].passed(ubuf & 1):
                yield
            assert self.acc[buf] == tile, 'epilogue read the wrong accumulator'
            if self.rng.random() < 0.7:
                yield                                    # the epilogue takes a while
            self.acc[buf] = None
            self.done_tiles.append(tile)
            self.acc_empty[buf].arrive()
            yield


def run(sa, sb, n_tiles, gather, seed):
    rng = random.Random(seed)
    tiles = list(range(n_tiles))
    chunks = {t: rng.randint(1, 2 * sa + 3) for t in tiles}        # active chunks per tile (>= 1, as the kernel guarantees)
    m = Model(sa, sb, tiles, chunks, gather, rng)
    agents = {'producer': m.producer(), 'mma': m.mma(), 'epilogue': m.epilogue()}
    idle = 0
    while agents:
        name = rng.choice(list(agents) + ['retire'])
        if name == 'retire':
            before = len(m.inflight)
            m.retire_one()
            progressed = len(m.inflight) != before
        else:
            snapshot = (tuple(b.phase for b in m.a_full + m.a_empty + m.b_full + m.b_empty + m.acc_full + m.acc_empty +
                              m.list_full + m.list_empty), len(m.inflight), len(m.done_tiles))
            try:
                next(agents[name])
            except StopIteration:
                del agents[name]
            progressed = snapshot != (tuple(b.phase for b in m.a_full + m.a_empty + m.b_full + m.b_empty + m.acc_full +
                                            m.acc_empty + m.list_full + m.list_empty), len(m.inflight), len(m.done_tiles))
        idle = 0 if progressed else idle + 1
        assert idle < 20000, 'deadlock: no role made progress (%s still running)' % sorted(agents)
    while m.inflight:
        m.retire_one()
    assert m.done_tiles == tiles


@pytest.mark.parametrize('sa,sb', [(10, 3), (9, 3), (8, 3), (6, 3), (2, 2), (3, 1)])
@pytest.mark.parametrize('gather', [True, False])
def test_v2_barrier_protocol_has_no_deadlock_or_hazard(sa, sb, gather):
    for seed in range(40):
        run(sa, sb, n_tiles=1 + seed % 7, gather=gather, seed=seed)
