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


def main():
    n = 0
    for nstage in (2, 3, 4):
        for tiles in (1, 2, 3, 4, 5, 8):
            for nkb in (1, 2, 3, 9):
                for seed in range(5):
                    simulate(nstage, tiles, nkb, seed)
                    n += 1
    print(f"persistent-kernel barrier protocol OK ({n} schedules)")


if __name__ == "__main__":
    main()
