"""Tree sparsification (apply_tree_filter_to_paf, src/tree_filter.rs:205-283), CPU side: the oracle's restatement
against hand-computed answers (the reference only tests extract_genome_prefix — parity unpinned), and the SipHash-1-3
used for the pseudo-random pairs against CPython's own siphash13 (zero key)."""
import os
import subprocess
import sys

import oracle_lib
import sweepga_b200 as swg

# siphash13, k0 = k1 = 0: `PYTHONHASHSEED=0 python -c 'hash(bytes) & (2**64 - 1)'` (sys.hash_info.algorithm == "siphash13")
SIPHASH13_VECTORS = {
    b"a": 4644417185603328019,
    b"abc": 13851880170939887858,
    b"HG002#1#\xffCHM13#1#\xff": 6036425489358445785,
    bytes(range(15)): 17514137373579004394,
    bytes(range(64)): 8493894268803903686,
}


def test_siphash13_matches_cpython():
    for msg, want in SIPHASH13_VECTORS.items():
        assert oracle_lib.siphash13(msg) == want, msg
    if sys.hash_info.algorithm == "siphash13":  # live cross-check on a few more messages
        msgs = [b"x" * k for k in (1, 7, 8, 9, 16, 23, 100)]
        code = "import sys; print([hash(bytes.fromhex(h)) & (2**64-1) for h in sys.argv[1:]])"
        out = subprocess.run([sys.executable, "-c", code] + [m.hex() for m in msgs], env={**os.environ, "PYTHONHASHSEED": "0"},
                             capture_output=True, text=True, check=True).stdout
        assert eval(out) == [oracle_lib.siphash13(m) for m in msgs]


def test_extract_genome_prefix_vectors():
    """src/tree_filter.rs:446-452 — the same rule as the scaffold sweep's chromosome-pair prefix."""
    assert swg.prefix_P2("HG002#1#chr1") == "HG002#1#"
    assert swg.prefix_P2("HG002#2#chr2") == "HG002#2#"
    assert swg.prefix_P2("NA12878#1#chrX") == "NA12878#1#"
    assert swg.prefix_P2("simple") == "simple"


def line(q, t, m, b, extra=""):
    return f"{q}\t1000\t0\t100\t+\t{t}\t1000\t0\t100\t{m}\t{b}\t60{extra}"


TREE_LINES = [
    line("A#1#c1", "B#1#c1", 90, 100),            # A-B 0.9 (with the next line: 180 / 200)
    line("B#1#c2", "A#1#c1", 90, 100, "\ttp:A:P"),
    line("A#1#c1", "C#1#c1", 80, 100),            # A-C 0.8
    line("D#1#c1", "A#1#c9", 70, 100),            # A-D 0.7
    line("B#1#c1", "C#1#c1", 95, 100),            # B-C 0.95
    line("B#1#c1", "D#1#c1", 60, 100),            # B-D 0.6
    line("C#1#c1", "D#1#c1", 85, 100),            # C-D 0.85
    line("A#1#c1", "A#1#c2", 100, 100),           # same genome: never written
    "# comment",
    "",
    "short\tline",
    line("C#1#c1", "D#1#c2", "x", "y"),           # unparsable: matches 0, block 1 -> C-D becomes 85 / 101
]


def run(tmp_path, k, f=0, r=0.0):
    src, out = tmp_path / "t.paf", tmp_path / "o.paf"
    src.write_text("\n".join(TREE_LINES) + "\n")
    kept, sel = oracle_lib.tree_filter_paf(str(src), str(out), k, f, r)
    lines = out.read_text().split("\n")
    assert lines[-1] == "" and len(lines) - 1 == kept
    return lines[:-1], sel


def test_oracle_tree_filter_known_answers(tmp_path):
    L = TREE_LINES
    # nearest neighbour of A is B, of B is C, of C is B, of D is C (85 / 101 = 0.84 still beats 0.7 and 0.6)
    got, sel = run(tmp_path, 1)
    assert sel == 3 and got == [L[0], L[1], L[4], L[6], L[11]]
    # plus the farthest: A-D, B-D, C-A, D-B -> every pair
    got, sel = run(tmp_path, 1, 1)
    assert sel == 6 and got == [L[0], L[1], L[2], L[3], L[4], L[5], L[6], L[11]]
    got, sel = run(tmp_path, 0, 0, 0.0)
    assert sel == 0 and got == []
    got, sel = run(tmp_path, 0, 0, 1.0)          # threshold saturates at u64::MAX: every pair
    assert sel == 6
    # the pseudo-random pairs: exactly those whose hash is below the threshold
    frac = 0.4
    want = 0
    for a, b in (("A", "B"), ("A", "C"), ("A", "D"), ("B", "C"), ("B", "D"), ("C", "D")):
        h = oracle_lib.siphash13(f"{a}#1#".encode() + b"\xff" + f"{b}#1#".encode() + b"\xff")
        want += h <= int(frac * 18446744073709551616.0)
    got, sel = run(tmp_path, 0, 0, frac)
    assert sel == want
