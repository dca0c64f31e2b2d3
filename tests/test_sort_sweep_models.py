"""Claims behind two round-2 device algorithms, checked on the CPU in plain Python / numpy (no GPU, no product code):

csrc/group_sort.cuh — the record sort as a counting sort by group.  Whatever slot ranges the runs of a group receive
(the atomics race across warps), marking the groups with a descent and repairing them by MERGING their ascending pieces — final
position of a word = offset inside its piece + number of smaller words in every other piece — yields exactly the stable order by
(group key, secondary key, index) that the LSD passes produce (src/paf_filter.rs:761-777: IndexMap grouping + sort_by_key).

csrc/sweep_tree.cuh — the n = 1 plane sweep of a pile: item m is kept iff it is the best (score desc, start asc, index asc) of the
active set S(p) = {k : start_k <= p < end_k} at some event position p in [start_m, end_m) and never overlaps the best of such a p
by more than the threshold.  best(p) comes from a segment tree over the distinct event positions with range-min paint and
leaf-to-root queries; runs of equal best(p) decide every item.  Checked against the oracle's statement-level restatement of
plane_sweep_query (src/plane_sweep_exact.rs:268-352) on random piles with ties and zero-length intervals."""
import numpy as np
import pytest

import oracle_lib


# ---- group sort ------------------------------------------------------------------------------------------------------------
def group_sort_model(grp, sec, rng, dead=None):
    """grp, sec: per item; returns the item order.  Mirrors run scan -> slot ranges in RANDOM run order -> table scan -> scatter ->
    emit (mark descents) -> merge pieces (binary-search ranks) / sort (more than 16 pieces)."""
    n = len(grp)
    live = np.ones(n, bool) if dead is None else grp != dead
    heads = np.ones(n, bool)
    heads[1:] = grp[1:] != grp[:-1]
    run_of = np.cumsum(heads) - 1
    run_start = np.nonzero(heads)[0]
    run_len = np.diff(np.append(run_start, n))
    n_runs = len(run_start)
    count, base = {}, np.zeros(n_runs, np.int64)
    for r in rng.permutation(n_runs):  # arrival order of the atomics: arbitrary
        g = grp[run_start[r]]
        if dead is not None and g == dead:
            continue
        base[r] = count.get(g, 0)
        count[g] = base[r] + run_len[r]
    keys = sorted(count)
    start, acc = {}, 0
    for g in keys:
        start[g] = acc
        acc += count[g]
    ib = max(1, int(n - 1).bit_length())
    words = np.zeros(acc, np.uint64)
    for i in range(n):
        if live[i]:
            words[start[grp[i]] + base[run_of[i]] + (i - run_start[run_of[i]])] = (np.uint64(sec[i]) << np.uint64(ib)) | np.uint64(i)
    out = words.copy()
    n_merged = n_sorted = 0
    for g in keys:
        s, e = start[g], start[g] + count[g]
        w = words[s:e]
        desc = np.nonzero(w[1:] < w[:-1])[0] + 1
        if len(desc) == 0:
            continue
        pst = np.concatenate(([0], desc, [len(w)]))
        if len(pst) - 1 <= 16:
            n_merged += 1
            res = np.zeros(len(w), np.uint64)
            for j in range(len(pst) - 1):
                a, b = pst[j], pst[j + 1]
                for x in range(a, b):
                    rank = x - a
                    for k in range(len(pst) - 1):
                        if k != j:
                            rank += int(np.searchsorted(w[pst[k]:pst[k + 1]], w[x], side="left"))
                    res[rank] = w[x]
            out[s:e] = res
        else:
            n_sorted += 1
            out[s:e] = np.sort(w)
    return (out & np.uint64((1 << ib) - 1)).astype(np.int64), n_merged, n_sorted


@pytest.mark.parametrize("seed", range(12))
def test_group_sort_model_equals_the_stable_sort(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 1500))
    n_groups = int(rng.integers(1, 12))
    # aligner-like rows: long runs of one group, interrupted by strays; inside a group mostly ascending, many ties
    grp = np.repeat(rng.integers(0, n_groups, n // 7 + 1), 7)[:n]
    stray = rng.random(n) < 0.15
    grp[stray] = rng.integers(0, n_groups, int(stray.sum()))
    sec = np.zeros(n, np.int64)
    for g in range(n_groups):
        ix = np.nonzero(grp == g)[0]
        vals = np.sort(rng.integers(0, 50, len(ix))) * 10
        wild = rng.random(len(ix)) < [0.0, 0.05, 0.5][seed % 3]
        vals[wild] = rng.integers(0, 500, int(wild.sum()))
        sec[ix] = vals
    dead = n_groups  # some excluded items
    grp2 = grp.copy()
    grp2[rng.random(n) < 0.05] = dead
    order, n_merged, n_sorted = group_sort_model(grp2, sec, rng, dead=dead)
    live = np.nonzero(grp2 != dead)[0]
    want = live[np.lexsort((live, sec[live], grp2[live]))]  # (group, secondary, index)
    assert np.array_equal(order, want)
    assert n_merged + n_sorted > 0  # the random arrival order of the runs' slot ranges did disorder some group


# ---- tree sweep --------------------------------------------------------------------------------------------------------------
def tree_sweep_model(start, end, score, thr):
    """n = 1 sweep of ONE group: items already in (start, index) order.  Returns keep[]."""
    n = len(start)
    live = end > start
    pos = np.unique(np.concatenate((start[live], end[live])))
    size = 1
    while size < max(len(pos), 1):
        size <<= 1
    order = np.lexsort((np.arange(n), start, -score))      # score desc, start asc, index asc
    rank = np.empty(n, np.int64)
    rank[order] = np.arange(n)
    NONE = n
    tree = np.full(2 * size, NONE, np.int64)
    lo, hi = np.searchsorted(pos, start), np.searchsorted(pos, end)
    for m in range(n):
        if not live[m]:
            continue
        l, r = lo[m] + size, hi[m] + size
        while l < r:
            if l & 1:
                tree[l] = min(tree[l], rank[m]); l += 1
            if r & 1:
                r -= 1; tree[r] = min(tree[r], rank[m])
            l >>= 1; r >>= 1
    best = np.full(len(pos), NONE, np.int64)
    for i in range(len(pos)):
        x = i + size
        while x >= 1:
            best[i] = min(best[i], tree[x]); x >>= 1
    keep = np.zeros(n, bool)
    for m in range(n):
        if not live[m]:
            continue
        good = flag = False
        for i in range(lo[m], hi[m]):
            b = best[i]
            if b == rank[m]:
                good = True
            elif b != NONE and thr < 1.0:
                o = order[b]
                ov = max(0, min(end[m], end[o]) - max(start[m], start[o]))
                ml = min(end[m] - start[m], end[o] - start[o])
                if ml > 0 and ov / ml > thr:
                    flag = True
        keep[m] = good and not flag
    return keep


@pytest.mark.parametrize("seed", range(10))
def test_tree_sweep_model_equals_the_reference_sweep(seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(2, 300))
    start = np.sort(rng.integers(0, 400, n))
    ln = rng.integers(0, 120, n)
    ln[rng.random(n) < 0.05] = 0
    end = start + ln
    identity = rng.choice([0.8, 0.9, 0.95, 1.0], n)
    thr = [0.0, 0.5, 0.95, 1.0][seed % 4]
    # the oracle's query sweep on one group: query interval = (start, end); scoring "ani" = identity (many ties)
    maps = [(int(start[i]), int(end[i]), 0, 1, float(identity[i])) for i in range(n)]
    kept = oracle_lib.plane_sweep("query", maps, 1, thr, scoring=0)
    want = np.zeros(n, bool)
    want[list(kept)] = True
    got = tree_sweep_model(start, end, identity, thr)
    assert np.array_equal(got, want), (np.nonzero(got != want)[0][:10], thr)
