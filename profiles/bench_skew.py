#!/usr/bin/env python
"""configs[4] (skew) at reduced scale: one dense pile + many tiny groups; times the GPU path and checks parity
against the oracle.  Usage: python profiles/bench_skew.py n_pile n_tiny_groups [check_oracle] [scaffold_dist]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sweepga_b200 as swg
from sweepga_b200 import synth
import oracle_lib
n_pile, n_tiny = int(sys.argv[1]), int(sys.argv[2])
check = len(sys.argv) > 3 and sys.argv[3] == "1"
t = synth.skew(n_pile=n_pile, n_tiny_groups=n_tiny, seed=5)
cfg = swg.FilterConfig.from_cli(scaffold_dist=sys.argv[4]) if len(sys.argv) > 4 else swg.FilterConfig()
ctx = swg.Context(0)
if n_pile < 10_000_000:  # two warm-ups (the call after an arena growth consolidates the blocks); skipped for the largest piles
    ctx.filter(cfg, t)
    ctx.filter(cfg, t)
t0 = time.time(); s, c, st = ctx.filter(cfg, t); dt = time.time() - t0
print(f"skew pile={n_pile} tiny={n_tiny}: n={t.n} gpu {dt*1e3:.1f} ms wall (device {st.ms_device:.1f} ms), kept {st.n_kept}, chains {st.n_chains_kept}", flush=True)
if check:
    t0 = time.time(); os_, oc, _ = oracle_lib.apply_filters(cfg, t); do = time.time() - t0
    print(f"oracle {do:.1f} s; identical status {np.array_equal(s, os_)} chain {np.array_equal(c, oc)}")
