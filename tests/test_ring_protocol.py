"""K7's stream ring (csrc/ag_solver.cuh, pgs_warp): a model of the consume / refill protocol under WORST-CASE arrival
of the asynchronous copies.  A piece that has been requested holds garbage until `cp.async.wait_group 3` forces it
complete (only what must be complete is complete); every record is read from the ring four trips after the last piece of
it was requested.  The model asserts that every float of every record read is the right stream float -- for random record
size sequences, for streams shorter than the ring, and for runs of the largest record."""
import numpy as np

R, PIECE, K, W = 1024, 32, 6, 3          # ring floats, floats per refill piece, pieces per record, groups in flight


def run(sizes, sweeps):
    total = sum(sizes)
    assert total % 32 == 0
    stream = np.arange(total)
    ring = np.array([stream[i % total] for i in range(R)])
    ppos, pabs = R % total, R
    cur = cabs = 0
    inflight = []
    minlead = 10 ** 9
    for _ in range(sweeps):
        for i, size in enumerate(sizes):
            at_end = cur + size >= total
            nxt = 0 if at_end else cur + size
            nabs = cabs + size
            while len(inflight) > W:                       # wait_group<W>
                for ri, sp in inflight.pop(0):
                    ring[ri:ri + PIECE] = stream[sp:sp + PIECE]
            for k in range(sizes[(i + 1) % len(sizes)]):   # header + lane blocks of the next record
                assert ring[(nabs + k) & (R - 1)] == nxt + k, (i, k)
            n = min(K, (nabs + R - pabs) // PIECE, (R - (pabs & (R - 1))) // PIECE, (total - ppos) // PIECE)
            grp = []
            for _j in range(n):
                ri = pabs & (R - 1)
                ring[ri:ri + PIECE] = -7                   # requested, not yet complete
                grp.append((ri, ppos))
                pabs += PIECE
                ppos += PIECE
            if ppos >= total:
                ppos = 0
            inflight.append(grp)
            minlead = min(minlead, pabs - nabs)
            cur, cabs = nxt, nabs
    return minlead


def _pad(sizes):
    return sizes + [16] if sum(sizes) % 32 else sizes


def test_random_record_sequences():
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 90))
        sizes = _pad([int(x) for x in rng.choice([48, 80, 112, 144], size=n, p=[.3, .4, .2, .1])])
        assert run(sizes, 6) >= 4 * 144 + 144


def test_worst_cases():
    for sizes in ([144] * 40, [144] * 7, [144, 144], [48], [80] * 3 + [16] + [112] * 3, [144] * 9 + [16] * 2):
        assert run(_pad(list(sizes)), 10) >= 4 * 144 + 144
