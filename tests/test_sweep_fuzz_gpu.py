"""The primitive sweeps (plane_sweep_query / target / both, src/plane_sweep_exact.rs:268-461) on single groups of every
pile shape against the oracle: sparse, deep piles (the thread-per-item kernel hands the group to the warp kernel), one
very long interval over many short ones (long leftward scans), equal starts / ends / scores, zero-length intervals, and
n = 1, 2, 3, unlimited with several overlap thresholds."""
import random

import pytest

import oracle_lib
import sweepga_b200 as swg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with swg.Context(0) as c:
        yield c


def make(rng, shape, n):
    out = []
    for i in range(n):
        if shape == "sparse":
            s = rng.randrange(0, 50 * n)
            ln = rng.randrange(0, 60)
        elif shape == "deep":
            s = rng.randrange(0, 2000)
            ln = rng.randrange(1, 3000)
        elif shape == "long_over_short":
            if i % 97 == 0:
                s, ln = rng.randrange(0, 1000), rng.randrange(20 * n, 60 * n)
            else:
                s, ln = rng.randrange(0, 60 * n), rng.randrange(1, 40)
        elif shape == "ties":
            s = rng.randrange(0, 40) * 10
            ln = rng.choice([0, 10, 10, 20, 30])
        else:  # nested
            c, h = rng.randrange(1000, 1100), rng.randrange(0, 1000)
            s, ln = c - h, 2 * h
        t = rng.randrange(0, 5000)
        tl = rng.randrange(0, 400) if shape != "ties" else rng.choice([0, 50, 100])
        idy = rng.choice([0.8, 0.9, 0.95, 0.99]) if shape == "ties" else rng.uniform(0.7, 1.0)
        out.append((s, s + ln, t, t + tl, idy))
    return out


@pytest.mark.parametrize("shape", ["sparse", "deep", "long_over_short", "ties", "nested"])
@pytest.mark.parametrize("seed", range(3))
def test_single_group_sweeps(ctx, shape, seed):
    rng = random.Random(1000 * seed + len(shape))
    for n in (2, 3, 17, 120, 700):
        m = make(rng, shape, n)
        for keep in (1, 2, 3, None):
            for thr in (0.95, 0.5, 0.0, 1.0):
                for scoring in (3, 4):
                    assert ctx.plane_sweep_query(m, keep, thr, scoring) == oracle_lib.plane_sweep("query", m, keep, thr, scoring), (shape, n, keep, thr, scoring)
                assert ctx.plane_sweep_target(m, keep, thr) == oracle_lib.plane_sweep("target", m, keep, thr), (shape, n, keep, thr)
            assert ctx.plane_sweep_both(m, keep, 1, 0.95) == oracle_lib.plane_sweep("both", m, keep, 0.95, n_keep2=1), (shape, n, keep)
