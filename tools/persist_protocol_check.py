"""Discrete simulation of the mbarrier protocol of csrc/tc_persist.cuh (no GPU): three roles (producer, MMA issuer,
epilogue x 4 warps) stepped in random order; checks that nothing deadlocks, that a stage is never overwritten before the
MMA that reads it has retired, and that an accumulator buffer is never overwritten before all four epilogue warps have
drained it.  mbarrier model: phase bit + pending count; wait(parity) passes when the current phase's parity != parity."""
import random


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase ^= 1
            self.pending = self.count

    def passed(self, parity):
        return self.phase != parity


def simulate(nstage, tiles, nkb, seed):
    rng = random.Random(seed)
    full = [MBar(1) for _ in range(nstage)]; empty = [MBar(1) for _ in range(nstage)]
    acc_full = [MBar(1), MBar(1)]; acc_empty = [MBar(4), MBar(4)]
    stage_owner = [None] * nstage          # (tile, kb) currently stored in the stage, None = free
    acc_state = [None, None]               # tile whose accumulator lives in the buffer; None = drained
    acc_readers = [0, 0]
    log = {"mma": [], "epi": []}

    def producer():
        it = 0
        for t in range(tiles):
            for kb in range(nkb):
                s, ph = it % nstage, (it // nstage) & 1
                while not empty[s].passed(ph ^ 1):
                    yield
                assert stage_owner[s] is None, "stage overwritten while in use"
                stage_owner[s] = (t, kb)
                full[s].arrive()           # TMA completes the expected bytes
                it += 1
                yield

    def issuer():
        it = 0
        for j in range(tiles):
            b, use = j & 1, j >> 1
            while not acc_empty[b].passed((use & 1) ^ 1):
                yield
            assert acc_state[b] is None and acc_readers[b] == 0, "accumulator overwritten before it was drained"
            acc_state[b] = j
            for kb in range(nkb):
                s, ph = it % nstage, (it // nstage) & 1
                while not full[s].passed(ph):
                    yield
                assert stage_owner[s] == (j, kb), f"MMA read the wrong stage contents {stage_owner[s]} != {(j, kb)}"
                log["mma"].append((j, kb))
                stage_owner[s] = None      # tcgen05.commit -> empty (the MMAs retired)
                empty[s].arrive()
                it += 1
                yield
            acc_readers[b] = 4
            acc_full[b].arrive()
            yield

    def epilogue(w):
        for j in range(tiles):
            b, use = j & 1, j >> 1
            while not acc_full[b].passed(use & 1):
                yield
            assert acc_state[b] == j, f"epilogue warp {w} read tile {acc_state[b]} instead of {j}"
            yield                           # tcgen05.ld ...
            log["epi"].append((w, j))
            acc_readers[b] -= 1
            if acc_readers[b] == 0:
                acc_state[b] = None
            acc_empty[b].arrive()
            yield

    roles = [producer(), issuer()] + [epilogue(w) for w in range(4)]
    alive = list(range(len(roles)))
    idle_rounds = 0
    while alive:
        i = rng.choice(alive)
        before = (len(log["mma"]), len(log["epi"]), [o for o in stage_owner])
        try:
            next(roles[i])
        except StopIteration:
            alive.remove(i)
        after = (len(log["mma"]), len(log["epi"]), [o for o in stage_owner])
        idle_rounds = 0 if before != after else idle_rounds + 1
        assert idle_rounds < 10000, "deadlock"
    assert log["mma"] == [(j, kb) for j in range(tiles) for kb in range(nkb)]
    assert sorted(log["epi"]) == sorted((w, j) for w in range(4) for j in range(tiles))


class TxBar:
    """full barrier of the multicast kernel: one arrival (expect_tx by the local producer) + (1 + CL) transfers; the
    transfers of peers may land before the local expect_tx (the pending arrival keeps the phase open)."""
    def __init__(self, units):
        self.units, self.arrived, self.tx, self.phase = units, 0, 0, 0

    def _maybe_flip(self):
        if self.arrived == 1 and self.tx == self.units:
            self.phase ^= 1
            self.arrived, self.tx = 0, 0

    def expect(self):
        assert self.arrived == 0
        self.arrived = 1
        self._maybe_flip()

    def complete(self):
        self.tx += 1
        assert self.tx <= self.units, "more transfers than one phase expects: a slice overtook a whole round"
        self._maybe_flip()

    def passed(self, parity):
        return self.phase != parity


def simulate_mcast(cl, nstage, nkb, seed):
    """csrc/tc_mcast.cuh: CL CTAs, each a producer and an MMA issuer; empty[s] of every CTA counts CL arrivals (every
    issuer's commit is multicast), full[s] = own A box + CL weight slices.  Checks: no deadlock, a stage is never
    written (by anybody's slice or the local A box) while the local MMAs of the previous round still read it."""
    rng = random.Random(seed)
    full = [[TxBar(1 + cl) for _ in range(nstage)] for _ in range(cl)]
    empty = [[MBar(cl) for _ in range(nstage)] for _ in range(cl)]
    reading = [[None] * nstage for _ in range(cl)]     # round whose data the stage holds / is being filled with
    consumed = [[-1] * nstage for _ in range(cl)]      # last round the local MMAs have retired from the stage
    mma_log = [[] for _ in range(cl)]

    def write_check(c, s, rnd):
        assert consumed[c][s] == rnd - nstage or rnd < nstage, \
            f"CTA {c} stage {s}: data of round {rnd} written before round {rnd - nstage} was consumed"

    def producer(c):
        for i in range(nkb):
            s, ph = i % nstage, (i // nstage) & 1
            while not empty[c][s].passed(ph ^ 1):
                yield
            full[c][s].expect()
            write_check(c, s, i)
            full[c][s].complete()                      # own A box
            yield
            for d in range(cl):                        # this CTA's slice, multicast to every CTA
                write_check(d, s, i)
                full[d][s].complete()
            yield

    def issuer(c):
        for i in range(nkb):
            s, ph = i % nstage, (i // nstage) & 1
            while not full[c][s].passed(ph):
                yield
            mma_log[c].append(i)
            consumed[c][s] = i
            for d in range(cl):                        # multicast commit
                empty[d][s].arrive()
            yield

    roles = [producer(c) for c in range(cl)] + [issuer(c) for c in range(cl)]
    alive = list(range(len(roles)))
    idle = 0
    while alive:
        i = rng.choice(alive)
        before = sum(len(m) for m in mma_log)
        try:
            next(roles[i])
        except StopIteration:
            alive.remove(i)
        idle = 0 if sum(len(m) for m in mma_log) != before else idle + 1
        assert idle < 20000, "deadlock"
    assert all(m == list(range(nkb)) for m in mma_log)


def main():
    n = 0
    for nstage in (2, 3, 4):
        for tiles in (1, 2, 3, 4, 5, 8):
            for nkb in (1, 2, 3, 9):
                for seed in range(5):
                    simulate(nstage, tiles, nkb, seed)
                    n += 1
    print(f"persistent-kernel barrier protocol OK ({n} schedules)")
    n = 0
    for cl in (2, 4):
        for nstage in (2, 3):
            for nkb in (1, 2, 5, 9, 50):
                for seed in range(5):
                    simulate_mcast(cl, nstage, nkb, seed)
                    n += 1
    print(f"multicast-kernel barrier protocol OK ({n} schedules)")


if __name__ == "__main__":
    main()
