#!/usr/bin/env python
"""Device-resident step time of the non-default modes on the configs[2] table (20 M PanSN mappings)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sweepga_b200 as swg
from sweepga_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20_000_000
t = synth.pansn(n, seed=3)
if "--shuffle" in sys.argv:  # the generator emits records grouped by genome pair and position-ordered, like an aligner; this
    # is the same table in random row order (sort scatter, gathers and first-appearance tables lose their locality)
    t = t.take(np.random.default_rng(11).permutation(t.n))
    print("rows shuffled")
n = t.n
ctx = swg.Context(0)
dev, dres = ctx.upload(t)
for name, flags in (("defaults", {}), ("rescue 100k", dict(scaffold_dist="100k")), ("1:1 / 1:1", dict(num_mappings="1:1", scaffold_filter="1:1")),
                    ("1:1 / 1:1 + rescue", dict(num_mappings="1:1", scaffold_filter="1:1", scaffold_dist="100k")),
                    ("many:many / 1:1", dict(scaffold_filter="1:1")), ("no scaffolding, 1:1", dict(num_mappings="1:1", scaffold_jump="0")),
                    ("scaffolds only", dict(scaffolds_only=True))):
    if "--only-defaults" in sys.argv and flags:
        continue
    cfg = swg.FilterConfig.from_cli(**flags)
    ctx.filter_device(cfg, dev, dres)
    ms = []
    for _ in range(3):
        st = ctx.filter_device(cfg, dev, dres)
        ms.append(st.ms_device)
    print(f"{name:22s} {min(ms):8.2f} ms  {n/min(ms)/1e3:8.1f} Mmappings/s  launches {st.gpu_launches}  kept {st.n_kept}  chains {st.n_chains_kept}  near-ties {st.score_near_ties}", flush=True)
