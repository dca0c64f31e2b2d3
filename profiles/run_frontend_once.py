#!/usr/bin/env python
"""One swg_filter_paf + one swg_ani_stats call on a synthetic PAF with cg:Z: and dv:f: tags (for ncu launch lists).
Usage: python profiles/run_frontend_once.py [n_lines]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sweepga_b200 as swg
from sweepga_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
t = synth.pansn(n, seed=3, n_hap=40, with_names=True)
d = tempfile.mkdtemp()
src = os.path.join(d, "in.paf")
names = np.array(t.names)
cols = [names[t.query_id], (t.query_end + 1000).astype(str), t.query_start.astype(str), t.query_end.astype(str),
        np.where(t.strand == ord("+"), "+", "-"), names[t.target_id], (t.target_end + 1000).astype(str), t.target_start.astype(str),
        t.target_end.astype(str), t.matches.astype(str), t.block_length.astype(str), np.full(t.n, "60")]
lines = cols[0]
for c in cols[1:]:
    lines = np.char.add(np.char.add(lines, "\t"), c)
lines = np.char.add(lines, np.char.add(np.char.add(np.char.add("\tdv:f:0.0", (t.matches % 97).astype(str)), "\tcg:Z:"),
                                        np.char.add(np.char.add(t.matches.astype(str), "="), np.char.add((t.block_length - t.matches).astype(str), "X"))))
with open(src, "w") as f:
    f.write("\n".join(lines.tolist()) + "\n")
ctx = swg.Context(0)
f = swg.PafFilter(swg.FilterConfig()); f._ctx = ctx
st = f.filter_paf(src, os.path.join(d, "out.paf"))
print("filter_paf:", st.gpu_launches, "launches, tokenise", round(st.ms_tokenize, 2), "ms, filter", round(st.ms_device, 2), "ms")
print("ani n100:", swg.ani_stats(ctx, src, "n100"))
