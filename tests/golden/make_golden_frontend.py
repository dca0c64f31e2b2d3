#!/usr/bin/env python
"""Regenerates the golden fixtures of the file front end: a small PAF with the parser quirks (frontend.paf), the
ORACLE's parse of it (frontend.npz) and the oracle's results for the callers around the filter (frontend.json:
filter_paf output digests, ANI pre-pass values as f64 bit patterns, tree-filter output digests).

Like make_golden.py these are oracle outputs (the Rust reference cannot be built here); they pin the oracle against
regressions and give the CUDA path an oracle-free target on the GPU box.  Usage: python tests/golden/make_golden_frontend.py
"""
import hashlib
import json
import os
import random
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sweepga_b200 as swg
from sweepga_b200 import synth
import oracle_lib

PAF = os.path.join(HERE, "frontend.paf")
FILTER_CASES = {"defaults": {}, "1to1": dict(num_mappings="1:1", scaffold_filter="1:1"), "rescue": dict(scaffold_dist="50k")}
ANI_CASES = ["all", "orthogonal", "n100", "n0.1-length", "n0.05-identity"]  # the last two cut inside the sorted list
TREE_CASES = [(1, 0, 0.0), (2, 1, 0.0), (0, 0, 0.3)]
COLS = ("query_id", "target_id", "query_start", "query_end", "target_start", "target_end", "block_length", "matches", "strand",
        "seq_genome_id", "seq_genome2_id")


def build_paf():
    rng = random.Random(7)
    t = synth.pansn(1500, seed=104, n_hap=5, with_names=True)
    lines = []
    for i in range(t.n):
        f = [t.names[t.query_id[i]], str(int(t.query_end[i]) + 1000), str(t.query_start[i]), str(t.query_end[i]), chr(t.strand[i]),
             t.names[t.target_id[i]], str(int(t.target_end[i]) + 2000), str(t.target_start[i]), str(t.target_end[i]),
             str(t.matches[i]), str(t.block_length[i]), "60"]
        r = i % 11
        if r in (0, 1, 2, 3):
            f += ["tp:A:P", "dv:f:%.*f" % (rng.randrange(2, 9), rng.random() * 0.2)]
        if r in (2, 5):
            f += ["cg:Z:%d=%dX" % (t.matches[i], t.block_length[i] - t.matches[i])]
        if r == 7:
            f += [rng.choice(["dv:f:1e-3", "dv:f:bad\tdv:f:0.05", "dv:f:0.1234567890123456789012", "cg:Z:=5", "cg:Z:5=X", "dv:f:-0.01", "cg:Z:0="])]
        if i % 97 == 0:
            f[9] = rng.choice(["x", "+40", "000000000000000000000000077", "1e3"])
        lines.append("\t".join(f) + ("\r" if i % 113 == 0 else ""))
        if i % 200 == 0:
            lines.append(rng.choice(["# header line", "", "too\tshort"]))
    return "\n".join(lines)  # no trailing newline


def digest(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def f64_bits(x):
    return "%016x" % struct.unpack("<Q", struct.pack("<d", x))[0]


def expected(tmpdir):
    """Everything the fixture holds, recomputed with the oracle."""
    t = oracle_lib.parse_paf(PAF)
    cols = {f: np.asarray(getattr(t, f)) for f in COLS}
    cols["identity_bits"] = t.identity.view(np.uint64)
    cols["rank"] = np.asarray(t.rank, np.uint64)
    js = {"names": t.names, "filter_paf": {}, "ani": {}, "tree": {}}
    out = os.path.join(tmpdir, "o.paf")
    for name, flags in FILTER_CASES.items():
        oracle_lib.filter_paf(swg.FilterConfig.from_cli(**flags), PAF, out)
        js["filter_paf"][name] = digest(out)
    for m in ANI_CASES:
        mm = swg.parse_ani_method(m)
        ani, pairs = oracle_lib.ani_stats(PAF, mm[0], mm[1], mm[2])
        js["ani"][m] = [f64_bits(ani), pairs]
    for k, f, r in TREE_CASES:
        kept, sel = oracle_lib.tree_filter_paf(PAF, out, k, f, r)
        js["tree"]["%d,%d,%g" % (k, f, r)] = [digest(out), kept, sel]
    return cols, js


if __name__ == "__main__":
    import tempfile
    with open(PAF, "w", newline="") as fh:
        fh.write(build_paf())
    with tempfile.TemporaryDirectory() as d:
        cols, js = expected(d)
    np.savez_compressed(os.path.join(HERE, "frontend.npz"), **cols)
    json.dump(js, open(os.path.join(HERE, "frontend.json"), "w"), indent=1)
    print("records", len(cols["rank"]), "names", len(js["names"]), js["ani"], js["tree"])
