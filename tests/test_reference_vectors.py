"""Known-answer tests transcribed from the reference's own test-suite (file:line cited per test).

Each test runs twice: against the CPU oracle (always — this is what PINS the oracle to the reference) and,
under `-m gpu`, against the CUDA path through the C ABI.  The assertions are the reference's assertions.
Excluded on purpose: tests/test_inter_chromosome_plane_sweep.rs:13-78 (stale, contradicts live
code); tests/test_genome_pair_grouping.rs:61-113 is used WITH --num-mappings 1:1 (it predates the many:many default).
"""
import os

import pytest

import oracle_lib
import sweepga_b200 as swg

MAX = None  # usize::MAX
LLI, IDENT, LENGTH, LENID = 3, 0, 1, 2


class OracleEngine:
    name = "oracle"

    def query(self, m, n, thr, sc=LLI):
        return oracle_lib.plane_sweep("query", m, n, thr, sc)

    def target(self, m, n, thr, sc=LLI):
        return oracle_lib.plane_sweep("target", m, n, thr, sc)

    def both(self, m, nq, nt, thr, sc=LLI):
        return oracle_lib.plane_sweep("both", m, nq, thr, sc, nt)

    def filter_paf(self, cfg, src, dst):
        oracle_lib.filter_paf(cfg, src, dst)


class GpuEngine:
    name = "gpu"

    def __init__(self):
        self.ctx = swg.Context(0)

    def query(self, m, n, thr, sc=LLI):
        return self.ctx.plane_sweep_query(m, n, thr, sc)

    def target(self, m, n, thr, sc=LLI):
        return self.ctx.plane_sweep_target(m, n, thr, sc)

    def both(self, m, nq, nt, thr, sc=LLI):
        return self.ctx.plane_sweep_both(m, nq, nt, thr, sc)

    def filter_paf(self, cfg, src, dst):
        f = swg.PafFilter(cfg)
        f._ctx = self.ctx
        f.filter_paf(src, dst)


@pytest.fixture(scope="module", params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def eng(request):
    return OracleEngine() if request.param == "oracle" else GpuEngine()


def run_paf(eng, tmp_path, text, **flags):
    src, dst = tmp_path / "in.paf", tmp_path / "out.paf"
    src.write_text(text)
    cfg = flags.pop("config", None) or swg.FilterConfig.from_cli(**flags)
    eng.filter_paf(cfg, str(src), str(dst))
    return [l for l in dst.read_text().split("\n") if l]


def mk(qs, qe, ts, te, identity=1.0):
    return (qs, qe, ts, te, identity)


# ---------------------------------------------------------------------------------------------
# src/plane_sweep_exact.rs:621-827 (in-file unit tests) and tests/test_plane_sweep.rs
# ---------------------------------------------------------------------------------------------
def test_empty_input(eng):  # plane_sweep_exact.rs:625-630, test_plane_sweep.rs:26-36
    if eng.name == "gpu":
        assert eng.query([], 1, 0.95) == []
    else:
        assert eng.query([], 1, 0.95) == []


def test_single_mapping(eng):  # plane_sweep_exact.rs:632-645, test_plane_sweep.rs:38-48
    assert eng.query([mk(100, 200, 300, 400, 0.95)], 1, 0.95) == [0]


def test_non_overlapping(eng):  # plane_sweep_exact.rs:647-673, test_plane_sweep.rs:50-71
    assert eng.query([mk(100, 200, 300, 400, 0.95), mk(300, 400, 500, 600, 0.90)], 1, 0.95) == [0, 1]


def test_overlapping_keep_best(eng):  # plane_sweep_exact.rs:675-702, test_plane_sweep.rs:73-95
    assert len(eng.query([mk(100, 200, 300, 400, 0.95), mk(150, 250, 350, 450, 0.90)], 1, 0.95)) == 2
    assert len(eng.query([mk(100, 250, 300, 450), mk(150, 350, 400, 600)], 1, 0.95)) == 2


def test_secondaries(eng):  # plane_sweep_exact.rs:704-741
    m = [mk(100, 200, 300, 400, 0.95), mk(100, 200, 500, 600, 0.90), mk(100, 200, 700, 800, 0.85)]
    kept = eng.query(m, 2, 0.95)
    assert kept == [0, 1]


def test_overlap_threshold_unit(eng):  # plane_sweep_exact.rs:743-797
    m = [mk(100, 200, 300, 400, 0.95), mk(100, 200, 500, 600, 0.90), mk(100, 200, 700, 800, 0.85)]
    assert len(eng.query(m, 1, 1.0)) == 1
    assert len(eng.query(m, 2, 1.0)) == 2
    assert len(eng.query(m, 2, 0.5)) == 2


def test_chromosome_boundaries(eng):  # plane_sweep_exact.rs:799-826 (u64::MAX coordinates)
    m = [mk(0, 100, 0, 100, 0.95), mk(2**64 - 101, 2**64 - 1, 1000, 1100, 0.90)]
    if eng.name == "oracle":
        assert len(eng.query(m, 1, 0.95)) == 2
    else:  # the device SoA is u32 by contract: the same extremes at the u32 boundary
        m = [mk(0, 100, 0, 100, 0.95), mk(2**32 - 101, 2**32 - 1, 1000, 1100, 0.90)]
        assert len(eng.query(m, 1, 0.95)) == 2


def test_identical_mappings(eng):  # test_plane_sweep.rs:97-135
    m = [mk(100, 200, 300, 400), mk(100, 200, 500, 600), mk(100, 200, 700, 800)]
    assert len(eng.query(m, 1, 0.95)) == 1
    assert len(eng.query(m, 2, 0.95)) == 2
    assert len(eng.query(m, MAX, 0.95)) == 3


def test_contained_mappings(eng):  # test_plane_sweep.rs:137-165
    m = [mk(100, 300, 400, 600), mk(150, 180, 500, 530)]
    assert eng.query(m, 1, 0.95) == [0]
    assert len(eng.query(m, 2, 0.95)) == 2


def test_overlap_threshold(eng):  # test_plane_sweep.rs:167-190
    m = [mk(100, 300, 400, 600), mk(100, 300, 700, 900), mk(100, 300, 1000, 1200), mk(100, 300, 1300, 1500)]
    assert len(eng.query(m, 2, 0.5)) == 2


def test_complex_overlaps(eng):  # test_plane_sweep.rs:192-215
    m = [mk(0, 100, 0, 100), mk(50, 150, 200, 300), mk(120, 220, 400, 500), mk(200, 300, 600, 700), mk(280, 380, 800, 900)]
    assert len(eng.query(m, 1, 0.95)) >= 3


def test_target_axis_filtering(eng):  # test_plane_sweep.rs:217-236
    m = [mk(100, 200, 300, 400), mk(300, 400, 350, 450), mk(500, 600, 600, 700)]
    assert 2 in eng.target(m, 1, 0.95)


def test_both_axes_filtering(eng):  # test_plane_sweep.rs:238-265
    m = [mk(100, 200, 300, 400), mk(100, 200, 500, 600), mk(300, 400, 300, 400), mk(500, 600, 700, 800)]
    assert 3 in eng.both(m, 1, 1, 0.95)


def test_score_calculation():  # test_plane_sweep.rs:267-293
    import math
    assert oracle_lib.score(100, 200, 1.0) > oracle_lib.score(100, 110, 1.0)
    r = oracle_lib.score(0, 1000, 1.0) / oracle_lib.score(0, 100, 1.0)
    assert abs(r - math.log(1000) / math.log(100)) < 0.001


def test_secondary_count(eng):  # test_plane_sweep.rs:295-337
    m = [mk(100, 200, 300, 400), mk(100, 190, 500, 590), mk(100, 180, 700, 780), mk(100, 170, 900, 970), mk(100, 160, 1100, 1160)]
    assert eng.query(m, 1, 1.0) == [0]
    assert len(eng.query(m, 3, 1.0)) == 3
    assert len(eng.query(m, MAX, 1.0)) == 5


def test_strand_independence(eng):  # test_plane_sweep.rs:339-356
    assert len(eng.query([mk(100, 200, 300, 400), mk(150, 250, 500, 600)], 1, 0.95)) == 2


def test_event_ordering(eng):  # test_plane_sweep.rs:358-376
    m = [mk(100, 100, 300, 300), mk(100, 200, 400, 500), mk(100, 300, 600, 800)]
    assert 0 not in eng.query(m, 1, 0.95)


def test_real_world_scenario(eng):  # test_plane_sweep.rs:378-424
    m = [mk(1000, 2000, 5000, 6000), mk(1500, 2500, 7000, 8000), mk(3000, 4000, 9000, 10000), mk(3200, 3800, 11000, 11600),
         mk(5000, 5500, 15000, 15500), mk(5000, 5500, 16000, 16500), mk(5000, 5500, 17000, 17500), mk(5000, 5500, 18000, 18500),
         mk(8000, 12000, 20000, 24000)]
    kept = eng.query(m, 1, 0.95)
    assert 8 in kept and len(kept) >= 4
    assert len(eng.query(m, 2, 0.95)) > len(kept)


# ---------------------------------------------------------------------------------------------
# tests/test_scoring_ranking.rs
# ---------------------------------------------------------------------------------------------
def test_identity_scoring_prefers_high_identity(eng):  # :25-40
    m = [mk(100, 500, 1000, 1400, 0.70), mk(100, 200, 2000, 2100, 0.99), mk(100, 300, 3000, 3200, 0.85)]
    assert eng.query(m, 1, 0.95, IDENT) == [1]


def test_length_scoring_prefers_long(eng):  # :42-57
    m = [mk(100, 200, 1000, 1100, 0.99), mk(100, 600, 2000, 2500, 0.50), mk(100, 350, 3000, 3250, 0.75)]
    assert eng.query(m, 1, 0.95, LENGTH) == [1]


def test_length_identity_scoring(eng):  # :59-74
    m = [mk(100, 200, 1000, 1100, 0.95), mk(100, 400, 2000, 2300, 0.60), mk(100, 300, 3000, 3200, 0.80)]
    assert eng.query(m, 1, 0.95, LENID) == [1]


def test_log_length_identity_scoring(eng):  # :76-91
    m = [mk(100, 200, 1000, 1100, 0.95), mk(100, 1100, 2000, 3000, 0.60), mk(100, 600, 3000, 3500, 0.75)]
    assert eng.query(m, 1, 0.95, LLI) == [2]


def test_ranking_with_identical_scores(eng):  # :93-109
    m = [mk(100, 300, 1000, 1200, 0.90), mk(100, 280, 2000, 2180, 1.00), mk(100, 460, 3000, 3360, 0.50)]
    assert len(eng.query(m, 1, 0.95, LENID)) == 1


def test_scoring_preserves_non_overlapping(eng):  # :111-122
    m = [mk(100, 200, 1000, 1100, 0.50), mk(300, 500, 2000, 2200, 0.99), mk(600, 700, 3000, 3100, 0.30)]
    assert len(eng.query(m, 1, 0.95, IDENT)) == 3


def test_overlapping_best_survives(eng):  # :124-139
    m = [mk(100, 300, 1000, 1200, 0.85), mk(150, 350, 2000, 2200, 0.90), mk(200, 400, 3000, 3200, 0.95)]
    assert 2 in eng.query(m, 1, 0.95, IDENT)


def test_scoring_with_contained(eng):  # :141-170
    m = [mk(100, 500, 1000, 1400, 0.80), mk(200, 300, 2000, 2100, 0.99)]
    assert 1 in eng.query(m, 1, 0.95, IDENT)
    assert 0 in eng.query(m, 1, 0.95, LENGTH)
    assert 0 in eng.query(m, 1, 0.95, LLI)


def test_ranking_order_multiple(eng):  # :172-195
    m = [mk(100, 200, 1000, 1100, 0.70), mk(100, 250, 2000, 2150, 0.80), mk(100, 300, 3000, 3200, 0.90),
         mk(100, 180, 4000, 4080, 0.99), mk(100, 220, 5000, 5120, 0.60)]
    kept = eng.query(m, 2, 0.95, LENID)
    assert sorted(kept) == [1, 2]


def test_extreme_values(eng):  # :197-241
    m = [mk(100, 101, 1000, 1001, 1.00), mk(100, 100100, 2000, 102000, 0.01), mk(100, 1100, 3000, 4000, 0.50)]
    assert eng.query(m, 1, 0.95, LENGTH)[0] == 1
    assert eng.query(m, 1, 0.95, IDENT)[0] == 0
    assert eng.query(m, 1, 0.95, LLI)[0] == 2


# ---------------------------------------------------------------------------------------------
# src/plane_sweep_scaffold.rs:292-371 — chains on one chromosome pair == plane_sweep_both(1,1)
# ---------------------------------------------------------------------------------------------
def test_scaffold_no_overlap(eng):  # :292-328
    assert len(eng.both([mk(0, 1000, 0, 1000, 0.95), mk(2000, 3000, 2000, 3000, 0.95)], 1, 1, 0.5)) == 2


def test_scaffold_overlapping_keeps_best(eng):  # :330-371
    kept = eng.both([mk(0, 1000, 0, 1000, 0.90), mk(900, 1900, 900, 1900, 0.98)], 1, 1, 0.95)
    assert 1 <= len(kept) <= 2
    if len(kept) == 1:
        assert kept == [1]


# ---------------------------------------------------------------------------------------------
# tests/test_plane_sweep_symmetry.rs — plane_sweep_core::plane_sweep (score = interval length)
# ---------------------------------------------------------------------------------------------
def _core(eng, spans, n, thr):
    iv = [(b, e, float(e - b)) for b, e in spans]
    if eng.name == "oracle":
        return oracle_lib.plane_sweep_core(iv, n, thr)
    return eng.ctx.plane_sweep_core(iv, n, thr)


def test_core_symmetry_simple(eng):  # :16-44
    m = [(100, 200, 300, 400), (150, 250, 350, 450), (300, 400, 100, 200)]
    assert len(_core(eng, [(a, b) for a, b, _, _ in m], 1, 0.95)) == 2
    assert len(_core(eng, [(c, d) for _, _, c, d in m], 1, 0.95)) == 2


def test_core_symmetry_transposed(eng):  # :46-80
    o = [(100, 500), (200, 400), (600, 900)]
    t = [(1000, 1400), (1100, 1300), (1500, 1800)]
    assert len(_core(eng, o, 2, 0.95)) == len(_core(eng, t, 2, 0.95))


def test_core_symmetry_with_overlaps(eng):  # :82-116
    m = [(0, 100), (50, 150), (200, 300), (250, 350)]
    for n in (1, 2, 3, 4):
        assert _core(eng, m, n, 0.95) == _core(eng, list(m), n, 0.95)


def test_core_asymmetric(eng):  # :118-152
    q = _core(eng, [(0, 200), (50, 150), (300, 500)], 1, 0.95)
    t = _core(eng, [(0, 100), (200, 400), (50, 150)], 1, 0.95)
    assert 2 in q and 1 in t and q != t


def test_core_perfect_symmetry(eng):  # :154-198
    m = [(100, 300), (400, 600), (700, 900)]
    for n in (1, 2, 3, MAX):
        k = _core(eng, m, n, 0.95)
        assert len(k) == 3


def test_core_gpu_matches_oracle_on_random_piles(eng):
    """not a reference test: the GPU secondary API against the oracle's restatement, incl. the greedy overlap pass"""
    if eng.name == "oracle":
        return
    import numpy as np
    rng = np.random.default_rng(5)
    for trial in range(12):
        n = int(rng.integers(2, 300))
        b = rng.integers(0, 2000, n)
        ln = rng.integers(0, 300, n)
        sc = rng.choice([1.0, 2.5, 7.0, 0.0, 3.25], n) * rng.integers(1, 4, n)
        iv = [(int(x), int(x + l), float(s)) for x, l, s in zip(b, ln, sc)]
        for keep, thr in ((1, 0.95), (2, 0.5), (3, 1.0), (None, 0.5), (1, 0.0)):
            assert eng.ctx.plane_sweep_core(iv, keep, thr) == oracle_lib.plane_sweep_core(iv, keep, thr), (trial, keep, thr)


# ---------------------------------------------------------------------------------------------
# pipeline-level vectors (PAF text + flags -> kept lines)
# ---------------------------------------------------------------------------------------------
def paf(*rows):
    return "".join("\t".join(str(x) for x in r) + "\n" for r in rows)


def test_default_plane_sweep(eng, tmp_path):  # tests/test_integration.rs:7-79
    text = paf(("query1", 1000, 100, 900, "+", "target1", 2000, 200, 1000, 800, 800, 60, "cg:Z:800M"),
               ("query1", 1000, 150, 850, "+", "target1", 2000, 300, 1000, 700, 700, 60, "cg:Z:700M"),
               ("query1", 1000, 200, 600, "+", "target1", 2000, 400, 800, 400, 400, 60, "cg:Z:400M"),
               ("query2", 1500, 100, 1400, "+", "target2", 2500, 100, 1400, 1300, 1300, 60, "cg:Z:1300M"),
               ("query2", 1500, 200, 1200, "+", "target2", 2500, 200, 1200, 1000, 1000, 60, "cg:Z:1000M"))
    out = run_paf(eng, tmp_path, text, scaffold_jump="0", num_mappings="1:1")
    assert len(out) == 2
    assert any("800\t800" in l for l in out) and any("1300\t1300" in l for l in out)
    assert all(l.endswith("\tst:Z:unassigned") and "ch:Z:" not in l for l in out)  # paf_filter.rs:409-434,1709-1718


def test_plane_sweep_with_secondaries(eng, tmp_path):  # tests/test_integration.rs:82-129
    text = paf(*[("chr1", 10000, 1000, 2000, "+", "chr1_ref", 10000, t, t + 1000, 1000, 1000, 60, "cg:Z:1000M")
                 for t in (1000, 3000, 5000, 7000, 9000)])
    assert len(run_paf(eng, tmp_path, text, num_mappings="3", scaffold_jump="0")) == 3


def test_plane_sweep_keep_all(eng, tmp_path):  # tests/test_integration.rs:132-180
    text = paf(("read1", 5000, 500, 1500, "+", "ref1", 10000, 1000, 2000, 1000, 1000, 60, "cg:Z:1000M"),
               ("read1", 5000, 1000, 1800, "+", "ref1", 10000, 2500, 3300, 800, 800, 60, "cg:Z:800M"),
               ("read1", 5000, 2000, 2600, "+", "ref1", 10000, 4000, 4600, 600, 600, 60, "cg:Z:600M"),
               ("read1", 5000, 3000, 3400, "+", "ref1", 10000, 5000, 5400, 400, 400, 60, "cg:Z:400M"))
    assert len(run_paf(eng, tmp_path, text, num_mappings="-1", scaffold_jump="0")) == 4


def test_plane_sweep_with_overlap_filtering(eng, tmp_path):  # tests/test_integration.rs:183-240
    text = paf(("contig1", 8000, 1000, 3000, "+", "ref1", 10000, 2000, 4000, 2000, 2000, 60, "cg:Z:2000M"),
               ("contig1", 8000, 1100, 2900, "+", "ref1", 10000, 5000, 6800, 1800, 1800, 60, "cg:Z:1800M"),
               ("contig1", 8000, 1200, 2800, "+", "ref1", 10000, 7000, 8600, 1600, 1600, 60, "cg:Z:1600M"),
               ("contig1", 8000, 4000, 5000, "+", "ref1", 10000, 4000, 5000, 1000, 1000, 60, "cg:Z:1000M"))
    out = run_paf(eng, tmp_path, text, num_mappings="1", overlap=0.5, scaffold_jump="0")
    assert len(out) >= 2 and any("4000\t5000" in l for l in out)


def test_mapping_plane_sweep_across_targets(eng, tmp_path):  # tests/test_mapping_plane_sweep.rs:8-54
    text = paf(("genome1#chrA", 100000, 10000, 20000, "+", "genome2#chrA", 100000, 10000, 20000, 9500, 10000, 60, "NM:i:500", "cg:Z:9500=500X"),
               ("genome1#chrA", 100000, 12000, 18000, "+", "genome2#chrB", 100000, 12000, 18000, 5400, 6000, 60, "NM:i:600", "cg:Z:5400=600X"))
    out = run_paf(eng, tmp_path, text, num_mappings="1:1", scaffold_jump="0", min_aln_identity="0", overlap=0.5)
    assert any("genome2#chrA" in l for l in out) and not any("genome2#chrB" in l for l in out)


def test_mapping_plane_sweep_target_axis(eng, tmp_path):  # tests/test_mapping_plane_sweep.rs:56-102
    text = paf(("genome1#chrA", 100000, 10000, 20000, "+", "genome2#chrX", 100000, 10000, 20000, 9500, 10000, 60, "NM:i:500", "cg:Z:9500=500X"),
               ("genome1#chrB", 100000, 10000, 20000, "+", "genome2#chrX", 100000, 12000, 22000, 9800, 10000, 60, "NM:i:200", "cg:Z:9800=200X"))
    out = run_paf(eng, tmp_path, text, num_mappings="1:1", scaffold_jump="0", min_aln_identity="0", overlap=0.5)
    assert any("genome1#chrB" in l for l in out) and not any("genome1#chrA" in l for l in out)


def _scaffold_rows(q, t, starts, tstarts, m, ln=5000):
    return [(q, 100000, s, s + ln, "+", t, 100000, ts, ts + ln, m, ln, 60, f"NM:i:{ln - m}", f"cg:Z:{m}={ln - m}X")
            for s, ts in zip(starts, tstarts)]


def test_scaffold_length_filtering(eng, tmp_path):  # tests/test_scaffold_length_filter.rs:7-75
    rows = [("query1", 100000, 10000 + i * 2000, 11000 + i * 2000, "+", "target", 100000, 10000 + i * 2000, 11000 + i * 2000, 950, 1000,
             60, "NM:i:50", "cg:Z:950=50X") for i in range(10)]
    rows += [("query2", 100000, 50000 + i * 2000, 51000 + i * 2000, "+", "target", 100000, 50000 + i * 2000, 51000 + i * 2000, 950,
              1000, 60, "NM:i:50", "cg:Z:950=50X") for i in range(5)]
    out = run_paf(eng, tmp_path, paf(*rows), scaffold_mass="10000", scaffold_jump="10000", min_aln_identity="0")
    assert len(out) == 10
    assert all(l.startswith("query1\t") for l in out)
    assert all(l.endswith("\tch:Z:chain_1\tst:Z:scaffold") for l in out)


def test_scaffold_span_not_mass(eng, tmp_path):  # tests/test_scaffold_length_filter.rs:78-126
    text = paf(("query", 150000, 0, 1000, "+", "target", 150000, 0, 1000, 950, 1000, 60, "NM:i:50", "cg:Z:950=50X"),
               ("query", 150000, 99000, 100000, "+", "target", 150000, 99000, 100000, 950, 1000, 60, "NM:i:50", "cg:Z:950=50X"))
    assert len(run_paf(eng, tmp_path, text, scaffold_mass="50000", scaffold_jump="100000", min_aln_identity="0")) == 2


def test_overlapping_scaffolds_same_chromosome_pair(eng, tmp_path):  # tests/test_scaffold_plane_sweep_filtering.rs:7-56
    rows = _scaffold_rows("chr1", "target_chr1", (10000, 15000), (10000, 15000), 4750) + \
        _scaffold_rows("chr1", "target_chr1", (12000, 17000), (30000, 35000), 4900)
    out = run_paf(eng, tmp_path, paf(*rows), scaffold_mass="1000", scaffold_jump="10000", min_aln_identity="0",
                  scaffold_filter="1:1", scaffold_dist="0")
    s = "\n".join(out)
    assert "12000\t17000" in s or "17000\t22000" in s
    assert not ("10000\t15000" in s or "15000\t20000" in s)
    assert len(out) == 2 and all(l.endswith("\tch:Z:chain_1\tst:Z:scaffold") for l in out)  # ranks 2,3 -> chain_1


def test_overlapping_scaffolds_different_targets(eng, tmp_path):  # tests/test_scaffold_plane_sweep_filtering.rs:59-118
    rows = _scaffold_rows("chr1", "target_chr1", (10000, 15000), (10000, 15000), 4750) + \
        _scaffold_rows("chr1", "target_chr2", (10000, 15000), (10000, 15000), 4900)
    s = "\n".join(run_paf(eng, tmp_path, paf(*rows), scaffold_mass="1000", scaffold_jump="10000", min_aln_identity="0", scaffold_filter="1:1"))
    assert "target_chr1" in s and "target_chr2" in s


def test_contained_scaffold_filtering(eng, tmp_path):  # tests/test_scaffold_plane_sweep_filtering.rs:121-169
    text = paf(("chr1", 100000, 15000, 18000, "+", "target_chr1", 100000, 15000, 18000, 2940, 3000, 60, "NM:i:60", "cg:Z:2940=60X"),
               ("chr1", 100000, 10000, 17500, "+", "target_chr1", 100000, 10000, 17500, 7125, 7500, 60, "NM:i:375", "cg:Z:7125=375X"),
               ("chr1", 100000, 17500, 25000, "+", "target_chr1", 100000, 17500, 25000, 7125, 7500, 60, "NM:i:375", "cg:Z:7125=375X"))
    s = "\n".join(run_paf(eng, tmp_path, text, scaffold_mass="1000", scaffold_jump="10000", min_aln_identity="0", scaffold_filter="1:1",
                          scaffold_dist="0"))
    assert "10000\t17500" in s or "17500\t25000" in s
    assert "15000\t18000" not in s


def test_scaffolds_on_different_query_chromosomes(eng, tmp_path):  # tests/test_scaffold_plane_sweep_filtering.rs:172-224
    rows = _scaffold_rows("query_chr1", "target_chr1", (10000, 15000), (10000, 15000), 4750) + \
        _scaffold_rows("query_chr2", "target_chr1", (10000, 15000), (10000, 15000), 4900)
    s = "\n".join(run_paf(eng, tmp_path, paf(*rows), scaffold_mass="1000", scaffold_jump="10000", min_aln_identity="0", scaffold_filter="1:1"))
    assert "query_chr1" in s and "query_chr2" in s


def test_plane_sweep_grouping_bug(eng, tmp_path):  # tests/test_grouping_bug.rs:7-96
    text = paf(("chrI_query", 10000, 1000, 2000, "+", "chrI_target1", 10000, 1000, 2000, 1000, 1000, 60, "cg:Z:1000M"),
               ("chrI_query", 10000, 1000, 2000, "+", "chrII_target2", 10000, 2000, 3000, 1000, 1000, 60, "cg:Z:1000M"),
               ("chrI_query", 10000, 1000, 2000, "+", "chrIII_target3", 10000, 3000, 4000, 1000, 1000, 60, "cg:Z:1000M"),
               ("chrII_query", 15000, 2000, 3000, "+", "chrI_target1", 10000, 2000, 3000, 1000, 1000, 60, "cg:Z:1000M"),
               ("chrII_query", 15000, 2000, 3000, "+", "chrII_target2", 10000, 4000, 5000, 1000, 1000, 60, "cg:Z:1000M"))
    out = run_paf(eng, tmp_path, text, num_mappings="1", scaffold_mass="0")
    assert sum(l.startswith("chrI_query") for l in out) == 3
    assert sum(l.startswith("chrII_query") for l in out) == 2
    assert len(out) == 5


def test_multi_target_filtering(eng, tmp_path):  # tests/test_grouping_bug.rs:99-158
    text = paf(*[("query1", 5000, 1000, 2000, "+", f"target_{c}", 10000, t, t + 1000, 1000, 1000, 60, "cg:Z:1000M")
                 for c, t in (("A", 3000), ("B", 5000), ("C", 7000), ("D", 1000))])
    assert len(run_paf(eng, tmp_path, text, num_mappings="1", scaffold_mass="0")) == 4


def test_plane_sweep_preserves_genome_pairs(eng, tmp_path):  # tests/test_genome_pair_grouping.rs:13-58
    text = paf(("A#1#chr1", 1000, 0, 500, "+", "B#1#chr1", 1000, 0, 500, 450, 500, 60, "cg:Z:500M"),
               ("A#1#chr1", 1000, 0, 500, "+", "C#1#chr1", 1000, 0, 500, 400, 500, 60, "cg:Z:500M"),
               ("A#1#chr1", 1000, 0, 500, "+", "D#1#chr1", 1000, 0, 500, 350, 500, 60, "cg:Z:500M"))
    out = run_paf(eng, tmp_path, text, scaffold_jump="0")
    s = "\n".join(out)
    assert len(out) == 3 and "B#1#chr1" in s and "C#1#chr1" in s and "D#1#chr1" in s


def test_plane_sweep_within_genome_pair(eng, tmp_path):  # tests/test_genome_pair_grouping.rs:61-113 (+ --num-mappings 1:1)
    text = paf(("A#1#chr1", 1000, 0, 500, "+", "B#1#chr1", 1000, 0, 500, 450, 500, 60, "cg:Z:500M"),
               ("A#1#chr1", 1000, 0, 500, "+", "B#1#chr2", 1000, 0, 500, 400, 500, 60, "cg:Z:500M"),
               ("A#1#chr2", 1000, 0, 500, "+", "B#1#chr1", 1000, 0, 500, 350, 500, 60, "cg:Z:500M"))
    out = run_paf(eng, tmp_path, text, scaffold_jump="0", num_mappings="1:1")
    assert len(out) == 1 and out[0].startswith("A#1#chr1\t1000\t0\t500\t+\tB#1#chr1")


def test_reverse_strand_scaffold_plane_sweep(eng, tmp_path):  # tests/test_centromere_plane_sweep.rs:20-82
    text = paf(("query", 250000000, 129142789, 132986703, "+", "target", 250000000, 129142789, 132986703, 2938926, 3843914, 60,
                "NM:i:904988", "cg:Z:2938926=904988X"),
               ("query", 250000000, 129213003, 137240549, "-", "target", 250000000, 131937578, 139967018, 6372479, 8027546, 60,
                "NM:i:1655067", "cg:Z:6372479=1655067X"))
    out = run_paf(eng, tmp_path, text, min_aln_identity="0", scaffold_jump="100000")
    assert sum("\t-\t" in l for l in out) > 0
    assert len(out) == 2  # both kept, chain_1 / chain_2
    assert out[0].endswith("\tch:Z:chain_1\tst:Z:scaffold") and out[1].endswith("\tch:Z:chain_2\tst:Z:scaffold")


def test_reverse_vs_forward_scaffold_scoring(eng, tmp_path):  # tests/test_centromere_plane_sweep.rs:84-129
    text = paf(("query", 100000000, 10000000, 11000000, "+", "target", 100000000, 10000000, 11000000, 950000, 1000000, 60, "NM:i:50000",
                "cg:Z:950000=50000X"),
               ("query", 100000000, 10000000, 12000000, "-", "target", 100000000, 20000000, 22000000, 1900000, 2000000, 60,
                "NM:i:100000", "cg:Z:1900000=100000X"))
    out = run_paf(eng, tmp_path, text, min_aln_identity="0", scaffold_jump="100000")
    assert any("\t-\t" in l for l in out)


def _chain_cfg():  # FilterConfig literal of tests/test_chaining_stability.rs:177-199
    return swg.FilterConfig(min_block_length=0, mapping_filter_mode=2, scaffold_filter_mode=2, overlap_threshold=0.0, scaffold_gap=10_000,
                            min_scaffold_length=0, scaffold_overlap_threshold=0.0, scaffold_max_deviation=20_000, scoring_function=3,
                            min_identity=0.0, min_scaffold_identity=0.0)


def _chains(lines):
    ch = {}
    for l in lines:
        f = l.split("\t")
        cid = [x for x in f if x.startswith("ch:Z:")]
        if cid:
            ch.setdefault(cid[0], []).append(f"{f[2]}-{f[3]}")
    return ch


def test_nearest_neighbor_chaining(eng, tmp_path):  # tests/test_chaining_stability.rs:148-246
    text = paf(("querySeq", 10000, 0, 1000, "+", "targetSeq", 10000, 0, 1000, 950, 1000, 60),
               ("querySeq", 10000, 1100, 2100, "+", "targetSeq", 10000, 1100, 2100, 950, 1000, 60),
               ("querySeq", 10000, 5000, 6000, "+", "targetSeq", 10000, 5000, 6000, 950, 1000, 60))
    ch = _chains(run_paf(eng, tmp_path, text, config=_chain_cfg()))
    assert len(ch) == 1
    members = next(iter(ch.values()))
    assert sorted(members) == ["0-1000", "1100-2100", "5000-6000"]


def test_overlap_penalty(eng, tmp_path):  # tests/test_chaining_stability.rs:248-350
    text = paf(("querySeq", 10000, 0, 1000, "+", "targetSeq", 10000, 0, 1000, 950, 1000, 60),
               ("querySeq", 10000, 900, 1900, "+", "targetSeq", 10000, 900, 1900, 950, 1000, 60),
               ("querySeq", 10000, 1100, 2100, "+", "targetSeq", 10000, 1100, 2100, 950, 1000, 60))
    ch = _chains(run_paf(eng, tmp_path, text, config=_chain_cfg()))
    assert ch
    a = [k for k, v in ch.items() if "0-1000" in v]
    c = [k for k, v in ch.items() if "1100-2100" in v]
    if a and c:
        assert a == c


# ---------------------------------------------------------------------------------------------
# tests/test_chain_monotonicity.rs:128-345 (CLI runs on synthetic PAFs; flags -> FilterConfig.from_cli)
# ---------------------------------------------------------------------------------------------
def _cg_row(qs, ident_pm, qlen=100000, strand="+", ts=None, ln=1000, name_q="query", name_t="target"):
    m = ln * ident_pm // 1000
    ts = qs if ts is None else ts
    return (name_q, qlen, qs, qs + ln, strand, name_t, qlen, ts, ts + ln, m, ln, 60, f"NM:i:{ln - m}", f"cg:Z:{m}={ln - m}X")


def test_simple_collinear_chaining(eng, tmp_path):  # tests/test_chain_monotonicity.rs:128-164 (PAF :19-37)
    text = paf(*[_cg_row(qs, 950) for qs in (0, 2000, 8000, 20000, 50000)])
    for gap in (2_000, 10_000, 30_000, 100_000):
        out = run_paf(eng, tmp_path, text, scaffold_jump=str(gap), min_aln_identity="0.90", scaffold_mass="0")
        assert len(out) == 5, gap


def test_mixed_identity_chaining(eng, tmp_path):  # tests/test_chain_monotonicity.rs:166-208 (PAF :41-66)
    text = paf(*([_cg_row(qs, 980, 200000) for qs in (0, 2000, 5000, 8000, 11000)] +
                 [_cg_row(qs, 900, 200000) for qs in (50000, 80000, 120000, 160000, 195000)]))
    for gap, thr, expected in ((10_000, "0.95", 5), (100_000, "0.95", 0), (10_000, "0.85", 10), (100_000, "0.85", 10)):
        out = run_paf(eng, tmp_path, text, scaffold_jump=str(gap), min_scaffold_identity=thr, scaffold_mass="0")
        assert len(out) == expected, (gap, thr)


def test_fragmented_chaining_coverage(eng, tmp_path):  # tests/test_chain_monotonicity.rs:210-249 (PAF :70-93)
    text = paf(*[_cg_row(i * 3000, 950 + (i % 3) * 10) for i in range(20)])
    for gap in (5_000, 50_000, 500_000):
        assert len(run_paf(eng, tmp_path, text, scaffold_jump=str(gap), min_aln_identity="0.90", scaffold_mass="0")) == 20, gap


def test_centromere_inversion_filtering(eng, tmp_path):  # tests/test_chain_monotonicity.rs:251-345
    text = paf(*[_cg_row(qs, 760, 200000000, "-", ts, 1000000) for qs, ts in
                 ((129000000, 132000000), (130000000, 133000000), (131000000, 134000000))])
    flags = dict(scaffold_jump="10000", scaffold_mass="0")
    assert len(run_paf(eng, tmp_path, text, min_aln_identity="0.80", **flags)) == 0   # 76 % < 80 %
    assert len(run_paf(eng, tmp_path, text, min_aln_identity="0.75", **flags)) > 0    # 76 % >= 75 %
    assert len(run_paf(eng, tmp_path, text, min_aln_identity="0", **flags)) > 0


def test_chaining_monotonicity_property(eng, tmp_path):
    """tests/test_chaining_stability.rs:52-94 runs the aligner on data/scerevisiae8.fa.gz (absent here: FASTA blob and FastGA
    missing).  Its assertion — the number of chain members never decreases as --scaffold-jump grows — on the yeast-shaped
    synthetic PAF instead."""
    from sweepga_b200 import synth
    src = tmp_path / "y.paf"
    synth.write_paf(synth.yeast_like(6000, seed=33), str(src))
    text = src.read_text()
    counts = []
    for gap in (10_000, 50_000, 100_000, 500_000, 1_000_000):
        ch = _chains(run_paf(eng, tmp_path, text, scaffold_jump=str(gap), min_aln_identity="0"))
        counts.append(sum(len(v) for v in ch.values()))
    assert counts == sorted(counts) and counts[0] > 0, counts


# ---------------------------------------------------------------------------------------------
# tests/test_error_handling.rs:46-366 (the PAF-level cases; the CLI's own messages are out of scope)
# ---------------------------------------------------------------------------------------------
GOOD12 = "seq1\t1000\t{qs}\t{qe}\t{strand}\tseq2\t2000\t100\t300\t150\t200\t60\n"


def test_malformed_paf_lines(eng, tmp_path):  # :46-77 — "rejected or produce no output"
    assert run_paf(eng, tmp_path, "seq1\t100\t200\n") == []


def test_invalid_paf_numbers(eng, tmp_path):  # :80-112 — a non-numeric field parses as 0 (paf_filter.rs:308-317); no output
    assert run_paf(eng, tmp_path, GOOD12.format(qs="NOT_A_NUMBER", qe=200, strand="+")) == []


def test_missing_file_error(eng, tmp_path):  # :115-144
    with pytest.raises((IOError, swg.SwgError)):
        eng.filter_paf(swg.FilterConfig(), str(tmp_path / "this_file_definitely_does_not_exist_12345.paf"), str(tmp_path / "o.paf"))


def test_invalid_coordinate_ranges(eng, tmp_path):  # :149-180 — start > end "currently passes them through"
    """The reference neither validates nor crashes (release build: wrapping u64 arithmetic).  The u32 table cannot hold a
    wrapped span, so the drop-in states the difference: such a record is an error only if it SURVIVES the
    stage-1 retain, and then a clean one (SWG_ERR_RANGE, context still usable); if the retain drops it, nothing happens."""
    text = GOOD12.format(qs=500, qe=200, strand="+")
    assert run_paf(eng, tmp_path, text, min_aln_length="1k") == []
    if eng.name == "gpu":
        with pytest.raises(swg.SwgError) as e:
            run_paf(eng, tmp_path, text)
        assert e.value.code == -2 and "end < start" in str(e.value)
        assert run_paf(eng, tmp_path, GOOD12.format(qs=0, qe=200, strand="+"), scaffold_jump="0") != []
    else:
        run_paf(eng, tmp_path, text)  # must not crash


def test_unsupported_format(eng, tmp_path):  # :183-218 — binary garbage: no record, no output (the CLI's message is its own)
    src, dst = tmp_path / "binary.bin", tmp_path / "o.paf"
    src.write_bytes(bytes([0xFF, 0xFE, 0xFD, 0xFC, 0x00, 0x01]))
    eng.filter_paf(swg.FilterConfig(), str(src), str(dst))
    assert dst.read_bytes() == b""


def test_partial_valid_paf(eng, tmp_path):  # :221-263
    text = ("seq1\t100\t0\t50\t+\tseq2\t200\t0\t50\t40\t50\t60\n"
            "INVALID LINE WITH GARBAGE\n"
            "seq3\t300\t0\t100\t+\tseq4\t400\t0\t100\t90\t100\t60\n"
            "too\tfew\tfields\n"
            "seq5\t500\t0\t150\t+\tseq6\t600\t0\t150\t140\t150\t60\n")
    out = run_paf(eng, tmp_path, text, scaffold_jump="0")
    assert len(out) == 3 and all(l.endswith("st:Z:unassigned") for l in out)


def test_negative_coordinates(eng, tmp_path):  # :266-295 — "-100" does not parse as u64 -> 0; the 200 bp mapping is below the scaffold mass
    assert run_paf(eng, tmp_path, GOOD12.format(qs=-100, qe=200, strand="+")) == []


def test_overflow_coordinates(eng, tmp_path):  # :298-327 — the overflowing number sits in the (unused) length column
    assert run_paf(eng, tmp_path, "seq1\t999999999999999999999\t0\t100\t+\tseq2\t2000\t0\t100\t90\t100\t60\n") == []


def test_invalid_strand(eng, tmp_path):  # :330-366 — any strand other than "+" is reverse (paf_filter.rs:311); no crash
    out = run_paf(eng, tmp_path, GOOD12.format(qs=0, qe=100, strand="X"), scaffold_jump="0")
    assert len(out) == 1


def test_empty_file(eng, tmp_path):  # :11-43 — the CLI refuses an empty input itself; the filter writes an empty output
    assert run_paf(eng, tmp_path, "") == []


# ---------------------------------------------------------------------------------------------
# tests/unit_tests.rs:148-208
# ---------------------------------------------------------------------------------------------
def test_cigar_extended_format(eng, tmp_path):  # :148-188 — the example CIGARs through parse_cigar_counts (src/paf.rs:32-64)
    examples = {"10=": 10, "5=2X3=": 8, "10=5I10=": 20, "10=5D10=": 20, "3=1X2=1I4=1D": 9}
    text = paf(*[("q", 1000, 0, 100, "+", "t", 1000, 0, 100, 1, 100, 60, f"cg:Z:{cg}") for cg in examples])
    src = tmp_path / "cg.paf"
    src.write_text(text)
    t = oracle_lib.parse_paf(str(src)) if eng.name == "oracle" else swg.parse_paf(str(src), eng.ctx)
    assert [int(x) for x in t.matches] == list(examples.values())
    assert t.identity.tolist() == [m / 100 for m in examples.values()]


def test_scaffold_annotations(eng, tmp_path):  # :190-208 — ch:Z:chain_<k> / st:Z:<status> as written by write_filtered_output
    text = paf(*[_cg_row(qs, 950) for qs in (0, 12000, 24000)], _cg_row(40000, 950, strand="-", ts=39000))
    out = run_paf(eng, tmp_path, text, scaffold_dist="100k")
    assert out
    for l in out:
        tags = [x for x in l.split("\t")[12:] if x[:5] in ("ch:Z:", "st:Z:")]
        assert len(tags) == 2
        for ann in tags:
            parts = ann.split(":")
            assert len(parts) == 3 and parts[0] in ("ch", "st") and parts[1] == "Z" and parts[2]
        assert tags[0].startswith("ch:Z:chain_") and tags[0][11:].isdigit()
        assert tags[1] in ("st:Z:scaffold", "st:Z:rescued")


# ---------------------------------------------------------------------------------------------
# src/pansn.rs:317-342, src/cli.rs parsers, src/main.rs:244-293 / src/library_api.rs:31-63 (host logic; no GPU)
# ---------------------------------------------------------------------------------------------
def test_round_nice():  # src/pansn.rs:300-315
    assert swg.round_nice(0) == 0
    assert swg.round_nice(950) == 1000 and swg.round_nice(2900) == 3000 and swg.round_nice(7200) == 7000
    assert swg.round_nice(10) == 50


def test_clamp_scaffold_params():  # src/pansn.rs:317-342
    assert swg.clamp_scaffold_params(50_000, 10_000, 1000, False) == (50_000, 10_000)
    assert swg.clamp_scaffold_params(50_000, 10_000, None, True) == (50_000, 10_000)
    assert swg.clamp_scaffold_params(50_000, 10_000, 1000, True) == (10_000, 600)
    assert swg.clamp_scaffold_params(5_000, 3_000, 1_000_000, True) == (5_000, 3_000)


def test_parse_metric_number():  # src/cli.rs:26-61
    assert swg.parse_metric_number("50k") == 50_000 and swg.parse_metric_number("10K") == 10_000
    assert swg.parse_metric_number("1.5m") == 1_500_000 and swg.parse_metric_number("2G") == 2_000_000_000
    assert swg.parse_metric_number("0") == 0 and swg.parse_metric_number("123") == 123
    for bad in ("", "abc", "5x", "k"):
        with pytest.raises(ValueError):
            swg.parse_metric_number(bad)


def test_parse_identity_value():  # src/cli.rs:76-130
    assert swg.parse_identity_value("0.9") == 0.9
    assert swg.parse_identity_value("90") == 0.9
    assert swg.parse_identity_value("1") == 1.0
    assert swg.parse_identity_value("ani50", 0.97) == 0.97
    assert abs(swg.parse_identity_value("ani50-2", 0.97) - 0.95) < 1e-12
    assert swg.parse_identity_value("ani50+10", 0.97) == 1.0
    with pytest.raises(ValueError):
        swg.parse_identity_value("ani50")
    with pytest.raises(ValueError):
        swg.parse_identity_value("bogus")


def test_parse_filter_mode_cli():  # src/main.rs:244-293
    f = swg.parse_filter_mode_cli
    assert f("1:1") == (0, 1, 1)
    assert f("1") == (1, 1, None) and f("1:many") == (1, 1, None) and f("1:∞") == (1, 1, None)
    assert f("many:1") == (2, None, 1) and f("∞:1") == (2, None, 1)
    for s in ("many:many", "many", "∞", "-1", "-1:-1", "MANY:MANY", "infinity"):
        assert f(s) == (2, None, None)
    assert f("10:5") == (2, 10, 5) and f("1:-1") == (1, 1, None) and f("0:3") == (2, None, 3)
    assert f("3") == (1, 3, None)
    assert f("a:b:c") == (0, 1, 1) and f("bogus") == (0, 1, 1)
    with pytest.raises(swg.SwgError):
        f("0")


def test_parse_filter_mode_library():  # src/library_api.rs:31-63 — differs from the CLI on "many:1"
    f = swg.parse_filter_mode
    assert f("many:many") == (2, None, None) and f("N:N") == (2, None, None)
    assert f("1:1") == (0, 1, 1)
    assert f("many:1") == (1, None, 1)
    assert f("1:many") == (1, 1, None)
    assert f("5:3") == (2, 5, 3) and f("5:many") == (2, 5, None)
    assert f("junk") == (0, 1, 1)
