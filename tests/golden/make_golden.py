#!/usr/bin/env python
"""Regenerates the committed golden fixtures: seeded synthetic tables -> ORACLE results (status, chain id).

The Rust reference cannot be built in this image, so these vectors are produced by the oracle, which is itself pinned
to the reference's own tests (tests/test_reference_vectors.py).  They serve as (a) a regression pin for the oracle and
(b) an oracle-free parity target for the CUDA path on the GPU box.  Usage: python tests/golden/make_golden.py
"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sweepga_b200 as swg
from sweepga_b200 import synth
import oracle_lib

CASES = {
    "yeast_defaults": (lambda: synth.yeast_like(6000, seed=101), {}),
    "yeast_1to1": (lambda: synth.yeast_like(6000, seed=101), dict(num_mappings="1:1", scaffold_filter="1:1")),
    "yeast_rescue": (lambda: synth.yeast_like(6000, seed=101), dict(scaffold_dist="100k")),
    "yeast_noscaffold_2": (lambda: synth.yeast_like(6000, seed=101), dict(num_mappings="2:2", scaffold_jump="0", overlap=0.5)),
    "pansn_defaults": (lambda: synth.pansn(20000, seed=102, n_hap=6), {}),
    "skew_rescue": (lambda: synth.skew(n_pile=4000, n_tiny_groups=800, seed=103, window=1_000_000), dict(scaffold_dist="20k")),
}

if __name__ == "__main__":
    for name, (make, flags) in CASES.items():
        t = make()
        cfg = swg.FilterConfig.from_cli(**flags)
        status, chain, st = oracle_lib.apply_filters(cfg, t)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), status=status, chain=chain,
                            table_checksum=np.array([int(t.query_start.astype(np.uint64).sum()), int(t.target_end.astype(np.uint64).sum()), t.n], np.uint64),
                            n_kept=np.array([st.n_kept, st.n_chains_kept], np.uint64))
        print(name, t.n, int(st.n_kept), int(st.n_chains_kept))
