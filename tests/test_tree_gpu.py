"""Tree sparsification on the device (swg_tree_filter_paf) against the oracle's restatement of
apply_tree_filter_to_paf (src/tree_filter.rs:205-283): the same output file, byte for byte."""
import random

import pytest

import oracle_lib
import sweepga_b200 as swg

from test_ani_gpu import pansn_paf
from test_tree_cpu import TREE_LINES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with swg.Context(0) as c:
        yield c


def both(ctx, src, tmp_path, k, f=0, r=0.0):
    a, b = tmp_path / "dev.paf", tmp_path / "orc.paf"
    got = swg.apply_tree_filter_to_paf(ctx, str(src), str(a), k, f, r)
    want = oracle_lib.tree_filter_paf(str(src), str(b), k, f, r)
    assert a.read_bytes() == b.read_bytes(), (k, f, r)
    assert got == want, (k, f, r)
    return got


def test_known_answers(ctx, tmp_path):
    src = tmp_path / "t.paf"
    src.write_text("\n".join(TREE_LINES) + "\n")
    assert both(ctx, src, tmp_path, 1) == (5, 3)
    assert both(ctx, src, tmp_path, 1, 1) == (8, 6)
    assert both(ctx, src, tmp_path, 0) == (0, 0)
    assert both(ctx, src, tmp_path, 0, 0, 1.0) == (8, 6)
    both(ctx, src, tmp_path, 0, 0, 0.4)
    both(ctx, src, tmp_path, 2, 0, 0.25)


@pytest.mark.parametrize("params", [(1, 0, 0.0), (2, 1, 0.0), (3, 0, 0.1), (0, 2, 0.5), (100, 0, 0.0)])
def test_synthetic_against_oracle(ctx, tmp_path, params):
    src = tmp_path / "a.paf"
    pansn_paf(str(src), 50000, seed=11)    # 8 haplotypes, dv / cg tags, '#' lines, CRLF, unparsable columns
    kept, sel = both(ctx, src, tmp_path, *params)
    assert sel > 0 and kept > 0


def test_edge_inputs(ctx, tmp_path):
    (tmp_path / "empty.paf").write_text("")
    assert both(ctx, tmp_path / "empty.paf", tmp_path, 1) == (0, 0)
    (tmp_path / "self.paf").write_text("A#1#x\t1\t0\t1\t+\tA#1#y\t1\t0\t1\t1\t1\t60\n# c\n")
    assert both(ctx, tmp_path / "self.paf", tmp_path, 1) == (0, 0)
    # huge integer columns (host patch), no trailing newline, no '#' in the names
    rng = random.Random(2)
    lines = ["g%d\t9\t0\t5\t+\tg%d\t9\t0\t5\t%s\t%s\t60" % (rng.randrange(6), rng.randrange(6), rng.choice(["5", "00000000000000000000007", "99999999999999999999999"]),
                                                      rng.choice(["10", "000000000000000000000010", ""])) for _ in range(3000)]
    (tmp_path / "big.paf").write_text("\n".join(lines))
    both(ctx, tmp_path / "big.paf", tmp_path, 1, 1)
    with pytest.raises(swg.SwgError):
        swg.apply_tree_filter_to_paf(ctx, str(tmp_path / "missing.paf"), str(tmp_path / "o.paf"), 1)
