"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, the host PAF
front end agrees with the oracle's restatement of extract_metadata, the shard planner, and the
"fail loudly without a device" contract."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib
import sweepga_b200 as swg
from sweepga_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sweepga_b200.h")).read()
    declared = set(re.findall(r"\b(swg_[a-z0-9_]+)\s*\(", header))
    declared -= {"swg_ctx", "swg_paf"}
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in include/sweepga_b200.h but not exported: {missing}"
    bound = {s[0] for s in _lib.SYMBOLS}
    assert declared <= bound, f"not bound in _lib.py: {sorted(declared - bound)}"


def _header_layout(structs):
    """sizeof / offsetof of every field of the given structs of include/sweepga_b200.h, as gcc lays them out."""
    import subprocess
    import tempfile
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sweepga_b200.h"', 'int main(void) {']
    for name, fields in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines += ['  return 0;', '}']
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return {k: int(v) for k, v in (l.split() for l in out.splitlines())}


def test_struct_layouts_match_header():
    """Every ctypes mirror (the product binding's and the oracle binding's own copies) has the header's size and offsets."""
    mirrors = {"swg_config": [_lib.swg_config, oracle_lib.swg_config], "swg_stats": [_lib.swg_stats, oracle_lib.swg_stats],
               "swg_mappings": [_lib.swg_mappings], "swg_result": [_lib.swg_result]}
    want = _header_layout({name: [f for f, _ in cls[0]._fields_] for name, cls in mirrors.items()})
    for name, classes in mirrors.items():
        for cls in classes:
            assert C.sizeof(cls) == want[name], (name, cls)
            for f, _ in cls._fields_:
                assert getattr(cls, f).offset == want[f"{name}.{f}"], (name, f, cls)


def test_config_default_matches_cli_defaults():
    c = _lib.swg_config()
    _lib.lib.swg_config_default(C.byref(c))
    d = swg.FilterConfig().to_c()
    for f, _ in _lib.swg_config._fields_:
        if f != "reserved":
            assert getattr(c, f) == getattr(d, f), f
    e = swg.FilterConfig.from_cli().to_c()
    for f, _ in _lib.swg_config._fields_:
        if f != "reserved":
            assert getattr(c, f) == getattr(e, f), f


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_device_fails_loudly():
    with pytest.raises(swg.SwgError) as e:
        swg.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


QUIRKY = "\n".join([
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60\ttp:A:P\tcg:Z:400=50X40I\tdv:f:0.1",   # dv after cg: dv wins
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60\tdv:f:0.1\tcg:Z:400=50X40I",            # cg after dv: cg wins
    "too\tshort",                                                                                       # consumes a rank
    "",
    "q#1#a\t1000\tx\t500\t-\tt#1#c\t2000\t20\t510\t450\tbad\t60",                                         # parse failures -> 0 / 1
    "q#1#a\t1000\t10\t500\t*\tt#1#b\t2000\t20\t510\t450\t490\t60\tcg:Z:490M",                              # '*' strand is reverse; M ignored
    "nohash\t1000\t10\t500\t+\ta#b\t2000\t20\t510\t+450\t490\t60\tcg:Z:=5",                                 # bad cigar ignored; '+450' parses
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60\tdv:f:abc\tcg:Z:0=",                      # bad dv ignored; 0 '=' ignored
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t0\t60\r",                                         # block 0, CRLF
]) + "\nq#1#a\t1000\t700\t900\t+\tt#1#b\t2000\t700\t900\t190\t200\t60"                                     # no trailing newline


def test_paf_parser_matches_oracle_on_quirks(tmp_path):
    p = tmp_path / "q.paf"
    p.write_text(QUIRKY)
    a, b = swg.parse_paf(str(p)), oracle_lib.parse_paf(str(p))
    assert a.n == b.n == 8
    assert list(a.rank) == list(b.rank) == [0, 1, 4, 5, 6, 7, 8, 9]
    assert a.names == b.names
    for f in ("query_id", "target_id", "query_start", "query_end", "target_start", "target_end", "block_length", "matches", "strand",
              "seq_genome_id", "seq_genome2_id"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(a.identity.view(np.uint64), b.identity.view(np.uint64))  # bit-exact f64
    assert a.identity[0] == 1.0 - 0.1 and a.identity[1] == 400 / 490
    assert a.query_start[2] == 0 and a.block_length[2] == 1 and chr(a.strand[2]) == "-" and chr(a.strand[3]) == "-"
    assert a.matches[4] == 450 and a.matches[1] == 400
    assert swg.prefix_P("q#1#a") == "q#1#" and swg.prefix_P2("a#b") == "a#b#" and swg.prefix_P("nohash") == "nohash"


def test_paf_parser_matches_oracle_on_synthetic(tmp_path):
    t = synth.yeast_like(40000, seed=2)   # > 1 MiB of text: exercises the multi-threaded chunking
    p = tmp_path / "y.paf"
    synth.write_paf(t, str(p))
    a, b = swg.parse_paf(str(p)), oracle_lib.parse_paf(str(p))
    assert a.n == b.n == t.n
    assert a.names == b.names
    for f in ("query_id", "target_id", "query_start", "query_end", "target_start", "target_end", "block_length", "matches", "strand",
              "seq_genome_id", "seq_genome2_id"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(a.identity.view(np.uint64), b.identity.view(np.uint64))
    assert np.array_equal(a.rank, np.arange(t.n, dtype=np.uint64))
    # the table round-trips: same coordinates and identities as generated
    assert np.array_equal(a.query_start, t.query_start) and np.array_equal(a.matches, t.matches)
    # ids are first-appearance ordered
    seen = []
    for q, tt in zip(a.query_id[:2000], a.target_id[:2000]):
        for x in (q, tt):
            if x not in seen:
                seen.append(int(x))
    assert seen == sorted(seen)


def test_paf_gz_and_bgz_inputs(tmp_path):
    """open_paf_input (src/paf.rs:10-28): extension gz / bgz => bgzf reader; bgzf is multi-member gzip."""
    import gzip
    t = synth.yeast_like(3000, seed=5)
    p = tmp_path / "a.paf"
    synth.write_paf(t, str(p))
    data = p.read_bytes()
    third = len(data) // 3
    blob = gzip.compress(data[:third]) + gzip.compress(data[third:2 * third]) + gzip.compress(data[2 * third:])
    (tmp_path / "a.paf.gz").write_bytes(blob)
    (tmp_path / "b.bgz").write_bytes(blob)
    a = swg.parse_paf(str(p))
    for name in ("a.paf.gz", "b.bgz"):
        b = swg.parse_paf(str(tmp_path / name))
        assert a.n == b.n and a.names == b.names
        assert np.array_equal(a.query_start, b.query_start) and np.array_equal(a.identity, b.identity)
    (tmp_path / "bad.gz").write_bytes(b"\x1f\x8b\x08\x00garbage-not-deflate")
    with pytest.raises(swg.SwgError):
        swg.parse_paf(str(tmp_path / "bad.gz"))


def test_paf_range_error(tmp_path):
    """A value beyond the u32 table (or end < start) does not reject the file at parse time: the record is kept with an
    impossible interval (query_start = 2^32 - 1 > query_end = 0) so that the filter can drop it in the stage-1 retain like
    the reference would (paf_filter.rs:384-388), and fails only if it survives (tests/test_parity_gpu.py)."""
    p = tmp_path / "big.paf"
    p.write_text("a\t1\t0\t5000000000\t+\tb\t1\t0\t10\t5\t10\t60\n"
                 "a\t1\t500\t200\t+\tb\t1\t0\t10\t5\t10\t60\n"
                 "a\t1\t0\t100\t+\tb\t1\t0\t10\t5\t10\t60\n")
    t = swg.parse_paf(str(p))
    assert t.n == 3
    assert (int(t.query_start[0]), int(t.query_end[0])) == (0xFFFFFFFF, 0)  # beyond u32: marked
    assert (int(t.query_start[1]), int(t.query_end[1])) == (500, 200)       # end < start: passed through
    assert (int(t.query_start[2]), int(t.query_end[2])) == (0, 100)
    assert t.identity[0] == 0.5
    with pytest.raises(swg.SwgError):
        swg.parse_paf(str(tmp_path / "missing.paf"))


def test_paf_writer_tags(tmp_path):
    p = tmp_path / "w.paf"
    p.write_text(QUIRKY)
    err = C.create_string_buffer(64)
    h = _lib.lib.swg_paf_parse(str(p).encode(), err, 64)
    n = _lib.lib.swg_paf_n_records(h)
    status = np.array([1, 0, 2, 3, 0, 0, 1, 1], np.uint8)
    chain = np.array([7, 0, 7, 0, 0, 0, 0, 12], np.uint32)
    out = tmp_path / "o.paf"
    assert _lib.lib.swg_paf_write(h, str(out).encode(), status.ctypes.data_as(_lib.u8p), chain.ctypes.data_as(_lib.u32p)) == 0
    _lib.lib.swg_paf_free(h)
    assert n == 8
    lines = out.read_text().split("\n")
    src = QUIRKY.split("\n")
    assert lines[0] == src[0] + "\tch:Z:chain_7\tst:Z:scaffold"
    assert lines[1] == src[4] + "\tch:Z:chain_7\tst:Z:rescued"
    assert lines[2] == src[5] + "\tst:Z:unassigned"
    assert lines[3] == src[8].rstrip("\r") + "\tst:Z:scaffold"
    assert lines[4] == src[9] + "\tch:Z:chain_12\tst:Z:scaffold"
    assert lines[5] == "" and len(lines) == 6


def test_shard_plan_keeps_genome_pairs_together_and_balances():
    t = synth.pansn(200_000, seed=9, n_hap=10)
    for k in (2, 4, 8):
        shard_of, sizes = swg.shard_plan(t, k)
        assert sizes.sum() == t.n and np.array_equal(np.bincount(shard_of, minlength=k), sizes.astype(np.int64))
        unit = t.seq_genome_id[t.query_id].astype(np.int64) * 1000 + t.seq_genome_id[t.target_id]
        for u in np.unique(unit)[:50]:
            assert np.unique(shard_of[unit == u]).size == 1
        assert sizes.max() <= 1.15 * sizes.mean()


def test_oracle_handles_u64_coordinates():
    """The oracle keeps the reference's u64 arithmetic (the device SoA is u32 by contract)."""
    kept = oracle_lib.plane_sweep("query", [(0, 100, 0, 100, 0.95), (2**64 - 101, 2**64 - 1, 1000, 1100, 0.9)], 1, 0.95)
    assert kept == [0, 1]


def test_oracle_plane_sweep_core_vectors():
    """tests/test_plane_sweep_symmetry.rs shape: plane_sweep_core keeps the n best at every Begin, then the
    greedy overlap pass (src/plane_sweep_core.rs:80-201)."""
    iv = [(100, 200, 0.9), (150, 250, 0.8), (300, 400, 0.7)]
    assert sorted(oracle_lib.plane_sweep_core(iv, 1, 0.95)) == [0, 2]
    assert sorted(oracle_lib.plane_sweep_core(iv, None, 0.95)) == [0, 1, 2]
    assert oracle_lib.plane_sweep_core([(0, 10, 1.0)], 1, 0.5) == [0]
    assert oracle_lib.plane_sweep_core([], 1, 0.5) == []


ANI_PAF = "\n".join([
    "A#1#c1\t1000\t0\t100\t+\tB#1#c1\t2000\t0\t100\t90\t100\t60",
    "A#1#c1\t1000\t0\t100\t+\tB#1#c1\t2000\t0\t200\t150\t200\t60",
    "A#1#c2\t500\t0\t100\t+\tC#1#c1\t3000\t0\t100\t50\t100\t60\ttp:A:P\tdv:f:bad\tdv:f:0.1\tdv:f:0.5",   # first dv that parses
    "A#1#c1\t1000\t0\t10\t+\tA#1#c2\t500\t0\t10\t10\t10\t60",                                            # same genome: skipped
    "# comment\twith\ttabs\t1\t2\t3\t4\t5\t6\t7\t8\t9",
    "",
    "short\tline",
]) + "\n"


def test_parse_ani_method_matches_reference_grammar():
    """parse_ani_method, src/main.rs:296-331."""
    P = swg.parse_ani_method
    assert P("all") == (swg.ANI_ALL, 0.0, swg.NSORT_IDENTITY) and P("ALL")[0] == swg.ANI_ALL
    assert P("orthogonal")[0] == swg.ANI_ORTHOGONAL and P("1:1")[0] == swg.ANI_ORTHOGONAL
    assert P("n50") == (swg.ANI_NPERCENTILE, 50.0, swg.NSORT_IDENTITY)
    assert P("N90-length") == (swg.ANI_NPERCENTILE, 90.0, swg.NSORT_LENGTH)
    assert P("n100-score") == (swg.ANI_NPERCENTILE, 100.0, swg.NSORT_SCORE)
    assert P("n12.5-identity-extra") == (swg.ANI_NPERCENTILE, 12.5, swg.NSORT_IDENTITY)
    for bad in ("", "n", "n0", "n101", "n-5", "n50-foo", "nan", "ninf", "median", "n50-"):
        assert P(bad) is None, bad


def test_oracle_ani_stats_known_answers(tmp_path):
    """calculate_ani_stats / calculate_ani_n_percentile (src/main.rs:334-688) on a hand-computed input (the reference has no
    test of its own for this function: parity unpinned, the expected values below follow the source line by line)."""
    p = tmp_path / "ani.paf"
    p.write_text(ANI_PAF)
    ab, ac = (90.0 + 150.0) / (100.0 + 200.0), ((1.0 - 0.1) * 100.0) / 100.0
    assert oracle_lib.ani_stats(str(p), 0) == ((ab + ac) / 2.0, 2)
    # genome size 1000 + 2000 + 500 + 3000 = 6500; n100 never reaches it: every alignment is used
    assert oracle_lib.ani_stats(str(p), 2, 100.0, 1) == ((ab + ac) / 2.0, 2)
    # n1: threshold 65, the first alignment of the sorted list crosses it
    assert oracle_lib.ani_stats(str(p), 2, 1.0, 1) == (0.9, 1)      # identity order: 0.9 (line 1), 0.9 (line 3), 0.75
    assert oracle_lib.ani_stats(str(p), 2, 1.0, 0) == (0.75, 1)     # length order: 200, 100, 100
    # n4: threshold 260 -> identity order takes lines 1, 3 and 2 (100 + 100 + 200 >= 260 at the third)
    assert oracle_lib.ani_stats(str(p), 2, 4.0, 1) == ((ab + ac) / 2.0, 2)
    # n3: threshold 195 -> lines 1 and 3: pairs (A,B) = 0.9, (A,C) = 0.9
    assert oracle_lib.ani_stats(str(p), 2, 3.0, 1) == ((0.9 + ac) / 2.0, 2)
    (tmp_path / "none.paf").write_text("A#1#x\t1\t0\t1\t+\tA#1#y\t1\t0\t1\t1\t1\t60\n")
    assert oracle_lib.ani_stats(str(tmp_path / "none.paf"), 0) == (0.0, 0)


def _stub_alntopaf(tmp_path, paf_text, fail=False):
    """A stand-in for FastGA's ALNtoPAF (not in this image): checks the argument convention of src/main.rs:750-757
    (`-x -T<threads> <file>`), writes a PAF to its standard output."""
    exe = tmp_path / "ALNtoPAF"
    paf = tmp_path / "stub_source.paf"
    paf.write_text(paf_text)
    exe.write_text("#!/bin/sh\n"
                   + ('exit 3\n' if fail else
                      '[ "$1" = "-x" ] || exit 4\ncase "$2" in -T[0-9]*) ;; *) exit 5;; esac\n[ -f "$3" ] || exit 6\n'
                      f'/bin/cat "{paf}"\n'))
    exe.chmod(0o755)
    return exe


def test_aln_to_paf_bridge(tmp_path, monkeypatch):
    """swg_aln_to_paf: the external-converter route for .1aln input (src/main.rs:737-770).  No converter -> UNSUPPORTED; a failing
    converter -> IO error and no output left behind; else its standard output, byte for byte.  (.1aln parity itself is unpinned: the
    reference holds no .1aln vector and the codec lives in fastga-rs.)"""
    import sweepga_b200 as swg
    from sweepga_b200 import _lib
    aln = tmp_path / "x.1aln"
    aln.write_bytes(b"1 3 aln\n")
    out = tmp_path / "x.paf"
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    monkeypatch.delenv("SWG_ALNTOPAF", raising=False)
    with pytest.raises(swg.SwgError) as e:
        swg.api.aln_to_paf(str(aln), str(out))
    assert e.value.code == _lib.ERR_UNSUPPORTED
    text = "q\t100\t0\t50\t+\tt\t100\t0\t50\t50\t50\t60\n"
    exe = _stub_alntopaf(tmp_path, text)
    monkeypatch.setenv("SWG_ALNTOPAF", str(exe))           # explicit path
    swg.api.aln_to_paf(str(aln), str(out), threads=3)
    assert out.read_text() == text
    out.unlink()
    monkeypatch.delenv("SWG_ALNTOPAF")
    monkeypatch.setenv("PATH", f"{tmp_path}:/usr/bin:/bin")  # found on PATH
    swg.api.aln_to_paf(str(aln), str(out))
    assert out.read_text() == text
    _stub_alntopaf(tmp_path, text, fail=True)
    with pytest.raises(swg.SwgError) as e:
        swg.api.aln_to_paf(str(aln), str(out))
    assert e.value.code == _lib.ERR_IO and not out.exists()


def test_c_example_builds_and_runs_host_side(tmp_path):
    """examples/filter_paf.c: the boundary from plain C99 (no C++ in the header), linked against the product library; its
    host-only mode needs no GPU, and without a device the filter mode fails loudly (no CPU path)."""
    exe = str(tmp_path / "filter_paf")
    libdir = os.path.join(ROOT, "sweepga_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "filter_paf.c"),
                    "-o", exe, os.path.join(libdir, "libsweepga_b200.so"), f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe, "--plan"], capture_output=True, text=True)
    assert r.returncode == 0 and "shard loads 110 100" in r.stdout and "default scaffold_gap 50000" in r.stdout
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage:" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, str(tmp_path / "in.paf"), str(tmp_path / "out.paf")], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stderr
