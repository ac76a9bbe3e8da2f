"""CPU checks of the index logic of the EXPERIMENTAL pair-symmetric BVE path
(lpm_v2_b200/csrc/sym_kernels.cuh, opt-in with lpm_set_bve_variant(200)).

The kernels themselves only run on a B200; what can be verified here is the schedule they
implement, restated in numpy exactly as the CUDA code indexes it:

  * the (target block, source chunk) triangle visits every ordered pair of distinct active
    particles exactly once -- diagonal tiles one-sided, tiles above the diagonal both ways --
    for any chunk length, block size and rank count;
  * the recursive-halving warp reduction leaves, for every source of a batch and every component,
    the sum over all 32 lanes in exactly one lane, and that lane is the one that issues the RED.
"""
import numpy as np
import pytest

TS = 256        # kTile


def coverage(nsrc, T, block, chunk_tiles, world=1):
    """count[c, j] = how often the pair (target c, source j) is accumulated (sym_bve_kernel)."""
    TB = block * T
    DT = TB // TS
    assert TB % TS == 0
    nsrc_pad = -(-nsrc // TS) * TS
    ntiles = nsrc_pad // TS
    nblocks = -(-nsrc_pad // TB)
    nchunks = -(-ntiles // chunk_tiles)
    count = np.zeros((nsrc, nsrc), dtype=np.int32)
    ctas = 0
    for rank in range(world):
        for bid in range(nblocks * nchunks):
            I, ck = bid % nblocks, bid // nblocks
            if world > 1 and I % world != rank:
                continue
            kdiag = I * DT
            k0 = ck * chunk_tiles
            k1 = min(k0 + chunk_tiles, ntiles)
            k0 = max(k0, kdiag)
            if k0 >= k1:
                continue
            ctas += 1
            c0, c1 = I * TB, min((I + 1) * TB, nsrc)        # targets past nsrc are null particles
            if c0 >= c1:
                continue
            for k in range(k0, k1):
                j0, j1 = k * TS, min((k + 1) * TS, nsrc)
                if j0 >= j1:
                    continue
                if k < kdiag + DT:                          # diag_tile: one-sided, self pair excluded
                    count[c0:c1, j0:j1] += 1
                    lo, hi = max(c0, j0), min(c1, j1)
                    idx = np.arange(lo, hi)
                    count[idx, idx] -= 1
                else:                                       # sym_tile: a[t] += P_j/d and cb[j] += P_t/d
                    count[c0:c1, j0:j1] += 1
                    count[j0:j1, c0:c1] += 1
    return count, ctas


@pytest.mark.parametrize("nsrc,T,block,chunk_tiles,world", [
    (1280, 4, 128, 1, 1), (1280, 4, 128, 3, 1), (1500, 4, 128, 2, 1), (1500, 8, 128, 2, 1),
    (2049, 4, 128, 4, 1), (300, 4, 128, 8, 1), (255, 8, 128, 1, 1),
    (1500, 4, 128, 2, 2), (2049, 4, 128, 3, 3), (3000, 8, 128, 2, 8),
])
def test_triangle_visits_every_ordered_pair_once(nsrc, T, block, chunk_tiles, world):
    count, ctas = coverage(nsrc, T, block, chunk_tiles, world)
    expect = np.ones((nsrc, nsrc), dtype=np.int32) - np.eye(nsrc, dtype=np.int32)
    assert np.array_equal(count, expect)
    assert ctas > 0


def shfl_xor(v, off):
    return v[np.arange(32) ^ off]


def reduce_red(cb):
    """sym_reduce_red<SB, NC>: cb[lane, s, a] -> list of (lane, source, component, value) REDs."""
    SB, NC = cb.shape[1], cb.shape[2]
    lane = np.arange(32)
    if SB == 8:
        levels, low = [(16, 4), (8, 2), (4, 1)], [2, 1]
    else:
        levels, low = [(16, 2), (8, 1)], [4, 2, 1]
    v = cb.copy()
    for off, half in levels:
        bit = (lane & off) != 0
        nxt = np.empty((32, half, NC))
        for k in range(half):
            for a in range(NC):
                lo, hi = v[:, k, a], v[:, k + half, a]
                keep = np.where(bit, hi, lo)
                send = np.where(bit, lo, hi)
                nxt[:, k, a] = keep + shfl_xor(send, off)
        v = nxt
    v = v[:, 0, :]
    for off in low:
        v = v + np.stack([shfl_xor(v[:, a], off) for a in range(NC)], axis=1)
    if SB == 8:
        sidx, q = (lane >> 2) & 7, lane & 3
    else:
        sidx, q = (lane >> 3) & 3, lane & 7
    return [(l, int(sidx[l]), int(q[l]), v[l, q[l]]) for l in range(32) if q[l] < NC]


@pytest.mark.parametrize("SB,NC", [(8, 3), (4, 3), (4, 2), (8, 2), (4, 1)])
def test_recursive_halving_reduction(SB, NC):
    rng = np.random.default_rng(7)
    cb = rng.integers(-1000, 1000, size=(32, SB, NC)).astype(np.float64)      # integers: sums are exact
    reds = reduce_red(cb)
    seen = {}
    for _, s, a, val in reds:
        assert (s, a) not in seen, "one RED per (source, component)"
        seen[(s, a)] = val
    assert len(seen) == SB * NC
    total = cb.sum(axis=0)
    for (s, a), val in seen.items():
        assert val == total[s, a]
