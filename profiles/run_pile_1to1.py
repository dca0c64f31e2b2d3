#!/usr/bin/env python
"""--num-mappings 1:1 (and 1:1 / 1:1) on the configs[4] pile at reduced scale: the plane sweeps on a deep pile."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sweepga_b200 as swg
from sweepga_b200 import synth
n_pile = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
t = synth.skew(n_pile=n_pile, n_tiny_groups=100_000, seed=5)
ctx = swg.Context(0)
dev, dres = ctx.upload(t)
for name, flags in (("1:1, no scaffolding", dict(num_mappings="1:1", scaffold_jump="0")), ("1:1 / 1:1", dict(num_mappings="1:1", scaffold_filter="1:1"))):
    cfg = swg.FilterConfig.from_cli(**flags)
    ctx.filter_device(cfg, dev, dres)
    st = ctx.filter_device(cfg, dev, dres)
    print(f"pile {n_pile} + 100000 tiny groups, {name}: {t.n} records, {st.ms_device:.1f} ms on device, launches {st.gpu_launches}, kept {st.n_kept}", flush=True)
