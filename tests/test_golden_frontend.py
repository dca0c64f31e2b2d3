"""Committed golden vectors of the file front end (tests/golden/frontend.{paf,npz,json}, produced by
tests/golden/make_golden_frontend.py): the oracle and the host parser must still reproduce them (CPU); the device
tokeniser, swg_filter_paf, the ANI pre-pass and the tree filter must match them without consulting the oracle (GPU)."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_frontend as G
import sweepga_b200 as swg

GOLD = np.load(os.path.join(HERE, "golden", "frontend.npz"))
JS = json.load(open(os.path.join(HERE, "golden", "frontend.json")))


def same_as_golden(t):
    assert t.names == JS["names"]
    assert np.array_equal(np.asarray(t.rank, np.uint64), GOLD["rank"])
    for f in G.COLS:
        assert np.array_equal(np.asarray(getattr(t, f)), GOLD[f]), f
    assert np.array_equal(t.identity.view(np.uint64), GOLD["identity_bits"])


def test_oracle_reproduces_frontend_golden(tmp_path):
    cols, js = G.expected(str(tmp_path))
    for k in GOLD.files:
        assert np.array_equal(cols[k], GOLD[k]), k
    assert js == JS


def test_host_parser_matches_frontend_golden():
    same_as_golden(swg.parse_paf(G.PAF))


@pytest.fixture(scope="module")
def ctx():
    with swg.Context(0) as c:
        yield c


@pytest.mark.gpu
def test_device_parser_matches_golden(ctx):
    same_as_golden(swg.parse_paf(G.PAF, ctx))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.FILTER_CASES))
@pytest.mark.parametrize("host", [False, True])
def test_filter_paf_matches_golden(ctx, tmp_path, name, host):
    f = swg.PafFilter(swg.FilterConfig.from_cli(**G.FILTER_CASES[name]))
    f._ctx = ctx
    out = tmp_path / "o.paf"
    f.filter_paf(G.PAF, str(out), host_frontend=host)
    assert G.digest(str(out)) == JS["filter_paf"][name]


@pytest.mark.gpu
@pytest.mark.parametrize("method", G.ANI_CASES)
def test_ani_matches_golden(ctx, method):
    ani, pairs = swg.ani_stats(ctx, G.PAF, method)
    assert [G.f64_bits(ani), pairs] == JS["ani"][method]


@pytest.mark.gpu
@pytest.mark.parametrize("case", G.TREE_CASES)
def test_tree_filter_matches_golden(ctx, tmp_path, case):
    out = tmp_path / "o.paf"
    kept, sel = swg.apply_tree_filter_to_paf(ctx, G.PAF, str(out), *case)
    assert [G.digest(str(out)), kept, sel] == JS["tree"]["%d,%d,%g" % case]
