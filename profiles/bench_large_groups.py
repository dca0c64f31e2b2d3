#!/usr/bin/env python
"""Chromosome-scale groups on ordinary (collinear) data: n_pairs chromosome pairs of two assemblies, each pair one collinear run of
`per_group` mappings (5 % of them on the other strand, 2 % displaced off the diagonal), so every (query, target, '+') group passes
FX_MIN_GROUP — the fixed point then runs on data without piles.  Compares it with the claims + sequential resolve
(SWG_NO_FIXPOINT=1).  Usage: python profiles/bench_large_groups.py per_group n_pairs"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sweepga_b200 as swg
from sweepga_b200 import synth
m, n_pairs = int(sys.argv[1]), int(sys.argv[2])
n_hap = n_pairs
rng = np.random.default_rng(11)
names = [f"A#1#chr{k + 1}" for k in range(n_pairs)] + [f"B#1#chr{k + 1}" for k in range(n_pairs)]
P, P2 = swg.prefix_ids(names)
cols = {k: [] for k in "qid tid qs qe ts te blen mat ident strand".split()}
for k in range(n_pairs):
    ln = rng.integers(1000, 8001, m)
    gap = rng.integers(0, 3000, m)
    qs = np.cumsum(ln + gap) - ln
    ts = qs + rng.integers(-300, 301, m) + np.where(rng.random(m) < 0.02, rng.integers(-200000, 200001, m), 0)
    ts = np.maximum(ts, 0)
    idn = rng.uniform(0.9, 0.999, m)
    cols["qid"].append(np.full(m, k)); cols["tid"].append(np.full(m, n_pairs + k))
    cols["qs"].append(qs); cols["qe"].append(qs + ln); cols["ts"].append(ts); cols["te"].append(ts + ln)
    cols["blen"].append(ln); cols["mat"].append(np.rint(idn * ln).astype(np.int64)); cols["ident"].append(np.rint(idn * ln) / ln)
    cols["strand"].append(np.where(rng.random(m) < 0.05, ord("-"), ord("+")).astype(np.uint8))
c = {k: np.concatenate(v) for k, v in cols.items()}
t = swg.MappingTable(c["qid"], c["tid"], c["qs"], c["qe"], c["ts"], c["te"], c["blen"], c["mat"], c["ident"], c["strand"], P, P2)
cfg = swg.FilterConfig()
ctx = swg.Context(0)
res = {}
for mode in ("fixpoint", "walk"):
    if mode == "walk":
        os.environ["SWG_NO_FIXPOINT"] = "1"
    for _ in range(2):
        ctx.filter(cfg, t)
    s, c, st = ctx.filter(cfg, t)
    res[mode] = (s, c)
    print(f"large groups n={t.n} n_hap={n_hap} {mode}: device {st.ms_device:.2f} ms, kept {st.n_kept}, chains {st.n_chains_kept}, launches {st.gpu_launches}", flush=True)
print("identical:", np.array_equal(res["fixpoint"][0], res["walk"][0]) and np.array_equal(res["fixpoint"][1], res["walk"][1]))
