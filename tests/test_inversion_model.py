"""The bucketed inversion capture of csrc/filter_pipeline.cu (inversion_grid), as a model in plain Python: a reverse-strand mapping goes
to the FIRST kept '+' chain (in chain_N order) of its chromosome pair whose query interval +- jump it touches and whose diagonal it is
within jump of (src/paf_filter.rs:553-596).  The grid enters every chain once per query bucket its extended interval touches, keyed
by (query bucket, diagonal bucket); a mapping looks at the query buckets it touches and at the diagonal buckets that
[dm - R, dm + R] touches, where R = ceil((jump + 1) * sqrt 2) + 1 bounds every deviation that can pass floor(|dev| / sqrt 2) <= jump.
Checked here: that bound, and that the grid finds the same chain as the walk over all chains, for every bucket width.  No GPU."""
import math

import numpy as np
import pytest

SQRT2 = 1.4142135623730951


def perp(dev):
    return int(dev / SQRT2)  # (u64)__ddiv_rn((double)dev, 1.4142135623730951)


@pytest.mark.parametrize("G", [0, 1, 7, 50, 1000, 50_000, 1_000_000, 2**31 - 2])
def test_deviation_bound(G):
    R = math.ceil((G + 1) * SQRT2) + 1
    assert perp(R + 1) > G and perp(R + 2) > G       # beyond R nothing passes
    lim = int((G + 1) * SQRT2)
    for dev in range(max(0, lim - 3), lim + 4):      # around the exact limit: passing implies dev <= R
        if perp(dev) <= G:
            assert dev <= R


def capture_walk(m, chains, G):
    qc, tc = (m[0] + m[1]) // 2, (m[2] + m[3]) // 2
    for u, (cqs, cqe, cts) in enumerate(chains):
        if m[1] < max(cqs - G, 0) or m[0] > cqe + G:
            continue
        if perp(abs((tc - qc) - (cts - cqs))) <= G:
            return u
    return None


def capture_grid(m, grid, G, wb, wd, doff):
    R = math.ceil((G + 1) * SQRT2) + 1
    qc, tc = (m[0] + m[1]) // 2, (m[2] + m[3]) // 2
    dm = tc - qc + doff
    best = None
    for b in range(m[0] >> wb, (m[1] >> wb) + 1):
        for dg in range(max(dm - R, 0) >> wd, ((dm + R) >> wd) + 1):
            for u, (cqs, cqe, cts) in grid.get((b, dg), []):  # ascending u inside a cell
                if best is not None and u >= best:
                    break
                if m[1] < max(cqs - G, 0) or m[0] > cqe + G:
                    continue
                if perp(abs((tc - qc) - (cts - cqs))) <= G:
                    best = u
                    break
    return best


@pytest.mark.parametrize("seed", range(9))
def test_grid_equals_the_walk_over_all_chains(seed):
    rng = np.random.default_rng(seed)
    G = int([20, 300, 5000][seed % 3])
    span = int([2_000, 60_000, 400_000][seed % 3])
    maxc = span + 10_000
    n_ch, n_m = int(rng.integers(1, 200)), 100
    cqs = rng.integers(0, span, n_ch)
    clen = rng.integers(1, [200, 5000, 60_000][seed % 3], n_ch)
    cts = np.where(rng.random(n_ch) < 0.6, cqs + rng.integers(-2 * G, 2 * G + 1, n_ch), rng.integers(0, span, n_ch)).clip(0)
    chains = [(int(a), int(a + l), int(t)) for a, l, t in zip(cqs, clen, cts)]
    if seed % 2:  # chain_N order of one (pair, '+') group is query order; the other seeds keep an arbitrary order
        chains.sort()
    mqs = rng.integers(0, span, n_m)
    mlen = rng.integers(1, 3000, n_m)
    mts = np.where(rng.random(n_m) < 0.6, mqs + rng.integers(-3 * G, 3 * G + 1, n_m), rng.integers(0, span, n_m)).clip(0)
    maps = [(int(a), int(a + l), int(t), int(t + l)) for a, l, t in zip(mqs, mlen, mts)]
    doff = maxc + 1
    for wb in (4, 13, 17):
        for wd in (3, 15, 30):
            grid = {}
            for u, (a, e, t) in enumerate(chains):
                dg = (t - a + doff) >> wd
                for b in range(max(a - G, 0) >> wb, (min(e + G, maxc) >> wb) + 1):
                    grid.setdefault((b, dg), []).append((u, (a, e, t)))
            for m in maps:
                assert capture_grid(m, grid, G, wb, wd, doff) == capture_walk(m, chains, G), (seed, wb, wd, m)
