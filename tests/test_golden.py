"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py): the oracle must still
reproduce them (CPU), and the CUDA path must match them without consulting the oracle (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden
import oracle_lib
import sweepga_b200 as swg


def _load(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    make, flags = make_golden.CASES[name]
    t = make()
    chk = np.array([int(t.query_start.astype(np.uint64).sum()), int(t.target_end.astype(np.uint64).sum()), t.n], np.uint64)
    assert np.array_equal(chk, g["table_checksum"]), "the seeded generator no longer reproduces the golden input"
    return t, swg.FilterConfig.from_cli(**flags), g


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    t, cfg, g = _load(name)
    s, c, _ = oracle_lib.apply_filters(cfg, t)
    assert np.array_equal(s, g["status"]) and np.array_equal(c, g["chain"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_gpu_matches_golden(ctx, name):
    t, cfg, g = _load(name)
    s, c, st = ctx.filter(cfg, t)
    assert np.array_equal(s, g["status"]) and np.array_equal(c, g["chain"])
    assert st.n_kept == int(g["n_kept"][0]) and st.n_chains_kept == int(g["n_kept"][1])
