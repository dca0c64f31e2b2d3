"""The claim behind csrc/chain_fixpoint.cuh, checked on the CPU in plain Python: the reference's sequential best-buddy loop
(src/paf_filter.rs:784-851) is the unique solution of

    pick(i) = first arg-min_j { d(i,j) : d(i,j) < B(i,j) },   B(i,j) = min { d(i',j) : i' < i, pick(i') = j }

and re-evaluating all positions against a snapshot of the picks (Jacobi rounds), starting from the unconstrained arg-min,
ends exactly there.  Also the re-evaluation test of k_fx_check: a position has to be re-evaluated iff its pick is now
blocked or a candidate ranking before its pick has become eligible.  No GPU, no product code: a model of the algorithm."""
import numpy as np
import pytest

NONE = -1


def gaps(a, b, fwd, G):
    """d(i,j) under the gap rule of paf_filter.rs:799-843, or None."""
    g5 = G // 5
    q = b[0] - a[1] if b[0] >= a[1] else (a[1] - b[0] if a[1] - b[0] <= g5 else G + 1)
    if fwd:
        r = b[2] - a[3] if b[2] >= a[3] else (a[3] - b[2] if a[3] - b[2] <= g5 else G + 1)
    else:
        r = a[2] - b[3] if a[2] >= b[3] else (b[3] - a[2] if b[3] - a[2] <= g5 else G + 1)
    return q * q + r * r if q <= G and r <= G else None


def candidates(rec, fwd, G):
    n = len(rec)
    out = []
    for i in range(n):
        c = []
        for j in range(i + 1, n):
            if rec[j][0] > rec[i][1] + G:
                break
            d = gaps(rec[i], rec[j], fwd, G)
            if d is not None:
                c.append((d, j))
        out.append(c)
    return out


def sequential(cands):
    """The reference's loop: strict '<' on both tests, first minimal j wins, the displaced predecessor gets no second pick."""
    n = len(cands)
    bps = [None] * n
    pick = [NONE] * n
    for i in range(n):
        best = None
        for d, j in cands[i]:
            if (best is None or d < best[0]) and (bps[j] is None or d < bps[j]):
                best = (d, j)
        if best:
            bps[best[1]] = best[0]
            pick[i] = best[1]
    return pick


def evaluate(i, cands, pickers):
    """arg-min over the candidates of i that are eligible against the snapshot `pickers[j]` = [(i', d')]."""
    best = None
    for d, j in cands[i]:
        blocked = any(ip < i and dp <= d for ip, dp in pickers[j])
        if not blocked and (best is None or (d, j) < best):
            best = (d, j)
    return best


def fixed_point(cands, exact_test):
    n = len(cands)
    cur = [min(c) if c else None for c in cands]  # unconstrained arg-min: smallest d, then smallest j
    rounds = evals = 0
    while True:
        pickers = [[] for _ in range(n)]
        for i, b in enumerate(cur):
            if b:
                pickers[b[1]].append((i, b[0]))
        todo = []
        for i in range(n):
            if exact_test:
                b = cur[i]
                need = b is not None and any(ip < i and dp <= b[0] for ip, dp in pickers[b[1]])
                if not need:  # X(i): the candidates ranking before the pick
                    for d, j in cands[i]:
                        if (b is None or (d, j) < b) and not any(ip < i and dp <= d for ip, dp in pickers[j]):
                            need = True
                            break
                if need:
                    todo.append(i)
            else:
                todo.append(i)
        changed = 0
        new = list(cur)
        for i in todo:
            new[i] = evaluate(i, cands, pickers)
            changed += new[i] != cur[i]
        evals += len(todo)
        if exact_test:
            assert changed == len(todo), "the re-evaluation test listed a position whose pick did not change"
        cur = new
        rounds += 1
        if changed == 0:
            return [b[1] if b else NONE for b in cur], rounds, evals
        assert rounds <= n + 1


def random_group(rng, n, span, max_len, jitter):
    qs = np.sort(rng.integers(0, span, n))
    ln = rng.integers(1, max_len + 1, n)
    diag = rng.random(n) < 0.6
    ts = np.where(diag, qs + rng.integers(-jitter, jitter + 1, n), rng.integers(0, span, n))
    ts = np.maximum(ts, 0)
    return [(int(a), int(a + l), int(t), int(t + l)) for a, l, t in zip(qs, ln, ts)]


@pytest.mark.parametrize("seed", range(24))
def test_jacobi_rounds_end_in_the_sequential_picks(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 220))
    rec = random_group(rng, n, span=[300, 2000, 20000][seed % 3], max_len=[20, 200][seed % 2], jitter=[3, 40][(seed // 2) % 2])
    G = [50, 400, 5000][seed % 3]
    for fwd in (True, False):
        c = candidates(rec, fwd, G)
        ref = sequential(c)
        full, rounds, _ = fixed_point(c, exact_test=False)
        assert full == ref
        lean, rounds2, evals = fixed_point(c, exact_test=True)
        assert lean == ref and rounds2 == rounds  # the exact test re-evaluates fewer positions, never different ones


def test_ties_and_duplicates():
    # identical records: every d ties, "first minimal j" decides; the chain 0 -> 1 -> 2 -> ... comes out of round 0
    rec = [(100, 110, 100, 110)] * 40
    c = candidates(rec, True, 50)
    assert fixed_point(c, True)[0] == sequential(c) == list(range(1, 40)) + [NONE]
    # a contested successor: the closer, later predecessor displaces the earlier one, which gets no second pick
    rec = [(0, 10, 0, 10), (2, 12, 2, 12), (30, 40, 30, 40), (31, 41, 200, 210)]
    c = candidates(rec, True, 50)
    assert fixed_point(c, True)[0] == sequential(c)
