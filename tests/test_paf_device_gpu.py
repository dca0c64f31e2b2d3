"""Device PAF front end (swg_paf_parse_device / swg_filter_paf): the GPU tokeniser and the GPU-assembled tagged
output against the oracle's restatement of extract_metadata / write_filtered_output (src/paf_filter.rs:292-376,
1689-1726, src/paf.rs:32-64), bit for bit, and against the host front end."""
import gzip
import random

import numpy as np
import pytest

import oracle_lib
import sweepga_b200 as swg
from sweepga_b200 import synth

pytestmark = pytest.mark.gpu

COLS = ("query_id", "target_id", "query_start", "query_end", "target_start", "target_end", "block_length", "matches", "strand",
        "seq_genome_id", "seq_genome2_id")


@pytest.fixture(scope="module")
def ctx():
    with swg.Context(0) as c:
        yield c


def same_table(a, b, what=""):
    assert a.n == b.n, what
    assert list(a.rank) == list(b.rank), what
    assert a.names == b.names, what
    for f in COLS:
        assert np.array_equal(getattr(a, f), getattr(b, f)), (what, f)
    assert np.array_equal(a.identity.view(np.uint64), b.identity.view(np.uint64)), (what, "identity bits")


def check(ctx, path, what=""):
    dev = swg.parse_paf(str(path), ctx)
    same_table(dev, oracle_lib.parse_paf(str(path)), what + " vs oracle")
    same_table(dev, swg.parse_paf(str(path)), what + " vs host front end")
    return dev


BASE = "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60"

NUMBER_QUIRKS = [
    "q#1#a\t1000\t+10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60",        # '+10' parses
    "q#1#a\t1000\t-10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60",        # '-10' -> default 0
    "q#1#a\t1000\t\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60",           # empty -> 0
    "q#1#a\t1000\t+\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60",          # '+' alone -> 0
    "q#1#a\t1000\t1 0\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60",        # inner space -> 0
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t\t60",            # empty block length -> 1
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t0\t0\t60",             # block length 0 -> max(.,1)
    "q#1#a\t1000\t000000000000000000000010\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t60",   # 24 digits, value 10 (host patch)
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t99999999999999999999999\t490\t60\tcg:Z:400=",  # overflow -> 0, cg rescues
    "q#1#a\t1000\t10\t500\t++\tt#1#b\t2000\t20\t510\t450\t490\t60",        # '++' is not '+'
    "q#1#a\t1000\t10\t500\t\tt#1#b\t2000\t20\t510\t450\t490\t60",          # empty strand
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490",             # exactly 11 fields, no tags
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450\t490\t",           # trailing tab
    "q#1#a\t1000\t10\t500\t+\tt#1#b\t2000\t20\t510\t450",                  # 10 fields: skipped
    "\t\t\t\t\t\t\t\t\t\t",                                                # 11 empty fields: a record of empty names
    "# comment",
    "#q\t1\t2\t3\t+\tt\t4\t5\t6\t7\t8\t9",                                 # '#' line with 12 fields is a record
]

DV = ["0.1", ".5", "5.", "1e-3", "1E-3", "1e+2", "-0.25", "+0.5", "0", "-0", "0.000", "1e", "e5", "0x10", "", ".", "-", "1.2.3",
      "inf", "-inf", "Infinity", "nan", "NaN", "infx", "0.1234567890123456789012", "1e-400", "1e400", "12345678901234567890",
      "9007199254740993", "0.30000000000000004", "1e22", "1e23", "123456789e-30", "4.9e-324", "0.1e1", "00000.5", "1_0", " 0.1",
      "0.1 ", "7e-1", "72e-2", "0.015625", "0.3333333333333333", "1e-22", "1e-23", "179769313486231570000e289"]

CG = ["10=5X3=", "=5", "5", "", "4000000000=", "18446744073709551616=", "99999999999999999999999=", "5=X", "0=", "0=0=",
      "00000000000000000000005=", "0000000000000000005=", "12M3=", "3=4", "3==", "3=\x01", "1=2X3I4D5=6N7S8H9P", "x", "5 =",
      "3=99999999999999999999999", "3=99999999999999999999999X", "7=  "]


def test_quirks_bit_exact(ctx, tmp_path):
    lines = list(NUMBER_QUIRKS)
    lines += [BASE + "\tdv:f:" + v for v in DV]
    lines += [BASE + "\tcg:Z:" + v for v in CG]
    lines += [BASE + "\tcg:Z:300=\tdv:f:" + v for v in DV[:12]]        # dv after cg wins when it parses
    lines += [BASE + "\tdv:f:0.5\tcg:Z:" + v for v in CG]               # cg after dv wins when it counts
    lines += [BASE + "\ttp:A:P\tNM:i:4\tdv:f:0.2\tzz:Z:cg:Z:5=", BASE + "\tdv:f:0.2\t", BASE + "\t\t\tdv:f:0.2", BASE + "\tdv:f",
              BASE + "\tcg:Z", BASE + "\tDV:F:0.2", BASE + "\r", BASE + "\tdv:f:0.2\r", ""]
    p = tmp_path / "q.paf"
    p.write_bytes("\n".join(lines).encode("latin-1") + b"\nq#1#a\t1000\t700\t900\t+\tt#1#b\t2000\t700\t900\t190\t200\t60")  # no final newline
    t = check(ctx, p, "quirks")
    assert t.n > 100


def rand_line(rng, names):
    q, t = rng.choice(names), rng.choice(names)
    qs = rng.randrange(0, 10 ** rng.randrange(1, 9))
    ts = rng.randrange(0, 10 ** rng.randrange(1, 9))
    ln = rng.randrange(0, 50000)
    bl = rng.randrange(0, 60000)
    mt = rng.randrange(0, bl + 1)
    f = [q, "100000000", str(qs), str(qs + ln), rng.choice("+-+-*"), t, "100000000", str(ts), str(ts + ln), str(mt), str(bl), "60"]
    for _ in range(rng.randrange(0, 5)):
        k = rng.randrange(6)
        if k == 0:
            f.append("dv:f:" + rng.choice(DV))
        elif k == 1:
            f.append("dv:f:%.*f" % (rng.randrange(0, 12), rng.random() * 0.3))
        elif k == 2:
            f.append("cg:Z:" + rng.choice(CG))
        elif k == 3:
            f.append("cg:Z:" + "".join(str(rng.randrange(1, 3000)) + rng.choice("=XID=M") for _ in range(rng.randrange(1, 40))))
        elif k == 4:
            f.append(rng.choice(["tp:A:P", "NM:i:12", "", "cm:i:5", "s1:i:100"]))
        else:
            f.append("dv:f:%g" % (rng.random() * 10 ** rng.randrange(-6, 1)))
    if rng.random() < 0.03:
        f = f[:rng.randrange(0, 11)]
    if rng.random() < 0.03 and len(f) > 10:
        f[2], f[3], f[7], f[8] = "5", "50", "5", "50"   # keep end >= start so that the line stays a record, not an error
        f[rng.choice([2, 7, 9, 10])] = rng.choice(["x", "", "+5", "1e3", "5.0", "-1"])
    return "\t".join(f) + ("\r" if rng.random() < 0.02 else "")


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("long_thresh", [None, "40"])
def test_fuzz_against_oracle(ctx, tmp_path, seed, long_thresh, monkeypatch):
    """long_thresh=40 sends practically every line through the warp-per-line kernel."""
    if long_thresh:
        monkeypatch.setenv("SWG_TOK_LONG", long_thresh)
    rng = random.Random(seed)
    names = ["g%d#%d#chr%d" % (rng.randrange(6), rng.randrange(2), rng.randrange(20)) for _ in range(60)] + ["plain", "a#b", "x#1#y#z", ""]
    lines = [rand_line(rng, names) for _ in range(4000)]
    for k in range(0, 4000, 211):
        lines.insert(k, "")
    p = tmp_path / "f.paf"
    p.write_bytes(("\n".join(lines) + ("\n" if seed % 2 else "")).encode("latin-1"))
    check(ctx, p, f"fuzz {seed} {long_thresh}")


def test_long_cigars(ctx, tmp_path):
    """Base-level CIGARs of tens of KB: warp-per-line path, dv before / after, Err in the middle."""
    rng = random.Random(3)
    lines = []
    for k in range(60):
        ops = "".join(str(rng.randrange(1, 5000)) + rng.choice("=X=I=D") for _ in range(rng.randrange(800, 12000)))
        if k % 7 == 3:
            ops = ops[:len(ops) // 2] + "X" + ops[len(ops) // 2:]   # may put two operations in a row -> Err
        if k % 11 == 5:
            ops += "18446744073709551616="                          # overflowing run -> Err (host patch on the warp path)
        tags = ["tp:A:P", "cg:Z:" + ops]
        if k % 3 == 0:
            tags.append("dv:f:0.0%d" % k)
        if k % 5 == 0:
            tags.insert(0, "dv:f:0.5")
        lines.append(BASE + "\t" + "\t".join(tags))
        lines.append(BASE)
    p = tmp_path / "long.paf"
    p.write_text("\n".join(lines) + "\n")
    check(ctx, p, "long cigars")


def test_synthetic_and_many_names(ctx, tmp_path):
    t = synth.yeast_like(60000, seed=4)
    p = tmp_path / "y.paf"
    synth.write_paf(t, str(p))
    d = check(ctx, p, "yeast")
    assert d.n == t.n and np.array_equal(d.query_start, t.query_start)
    # more distinct names than the first table size holds: the table is regrown
    rng = random.Random(1)
    lines = ["r%d\t1000\t0\t100\t+\tc%d#1#x\t5000\t%d\t%d\t90\t100\t60" % (rng.randrange(90000), i % 50, i, i + 100) for i in range(120000)]
    p2 = tmp_path / "reads.paf"
    p2.write_text("\n".join(lines) + "\n")
    d2 = check(ctx, p2, "many names")
    assert len(d2.names) > 40000


def test_compressed_empty_and_errors(ctx, tmp_path):
    t = synth.yeast_like(3000, seed=5)
    p = tmp_path / "a.paf"
    synth.write_paf(t, str(p))
    data = p.read_bytes()
    (tmp_path / "a.paf.gz").write_bytes(gzip.compress(data[:len(data) // 2]) + gzip.compress(data[len(data) // 2:]))
    same_table(swg.parse_paf(str(tmp_path / "a.paf.gz"), ctx), swg.parse_paf(str(p)), "gz")
    (tmp_path / "empty.paf").write_bytes(b"")
    assert swg.parse_paf(str(tmp_path / "empty.paf"), ctx).n == 0
    (tmp_path / "junk.paf").write_bytes(b"\n\nnot a paf line\n\n")
    assert swg.parse_paf(str(tmp_path / "junk.paf"), ctx).n == 0
    (tmp_path / "big.paf").write_text(BASE + "\na\t1\t0\t5000000000\t+\tb\t1\t0\t10\t5\t10\t60\n")
    # beyond u32 / end < start: kept as records with an impossible interval (the filter decides), same on both front ends
    same_table(swg.parse_paf(str(tmp_path / "big.paf"), ctx), swg.parse_paf(str(tmp_path / "big.paf")), "big")
    assert int(swg.parse_paf(str(tmp_path / "big.paf"), ctx).query_start[-1]) == 0xFFFFFFFF
    (tmp_path / "rev.paf").write_text("a\t1\t50\t10\t+\tb\t1\t0\t10\t5\t10\t60\n")
    same_table(swg.parse_paf(str(tmp_path / "rev.paf"), ctx), swg.parse_paf(str(tmp_path / "rev.paf")), "rev")
    with pytest.raises(swg.SwgError):
        swg.parse_paf(str(tmp_path / "missing.paf"), ctx)
    # the context is still usable after an error
    assert swg.parse_paf(str(p), ctx).n == t.n


@pytest.mark.parametrize("flags", [{}, dict(num_mappings="1:1", scaffold_filter="1:1"), dict(scaffold_dist="50k"), dict(scaffold_jump="0"),
                                   dict(num_mappings="many:many", scaffold_filter="many:many")])
def test_filter_paf_device_equals_host_and_oracle(ctx, tmp_path, flags):
    t = synth.yeast_like(20000, seed=12)
    src = tmp_path / "y.paf"
    synth.write_paf(t, str(src))
    lines = src.read_text().split("\n")
    rng = random.Random(7)
    for k in range(0, len(lines), 97):
        lines[k] = lines[k] + rng.choice(["\tcg:Z:100=2X50=", "\tdv:f:0.03", "\r", "\tdv:f:1e-2\tcg:Z:7=", ""])
    lines.insert(5, "short\tline")
    lines.insert(50, "")
    src.write_text("\n".join(lines))
    cfg = swg.FilterConfig.from_cli(**flags)
    f = swg.PafFilter(cfg)
    f._ctx = ctx
    a, b, c = tmp_path / "dev.paf", tmp_path / "host.paf", tmp_path / "orc.paf"
    st = f.filter_paf(str(src), str(a))
    f.filter_paf(str(src), str(b), host_frontend=True)
    oracle_lib.filter_paf(cfg, str(src), str(c))
    assert a.read_bytes() == c.read_bytes()
    assert b.read_bytes() == c.read_bytes()
    assert st.gpu_launches > 10 and st.ms_tokenize > 0


def test_filter_file_and_apply_paf_filter(tmp_path, monkeypatch):
    """unified_filter::filter_file (src/unified_filter.rs:280-347) and library_api::apply_paf_filter
    (src/library_api.rs:267-281): PAF input goes through filter_paf with the caller's keep_self / with keep_self = false;
    a ONEcode container ("1 " magic) goes through ALNtoPAF when there is one and PAF output is asked for, else it is reported as
    unsupported, never guessed."""
    import os
    from sweepga_b200 import _lib
    t = synth.yeast_like(6000, seed=21)
    src = tmp_path / "y.paf"
    synth.write_paf(t, str(src))
    # self mappings make keep_self observable
    lines = src.read_text().split("\n")
    extra = []
    for k in range(0, 200, 7):
        f = lines[k].split("\t")
        f[5] = f[0]
        extra.append("\t".join(f))
    src.write_text("\n".join(extra + lines))
    cfg = swg.FilterConfig.from_cli(scaffold_dist="20k")
    outs = {}
    for keep_self in (False, True):
        out, ref = tmp_path / f"o{int(keep_self)}.paf", tmp_path / f"r{int(keep_self)}.paf"
        swg.filter_file(str(src), str(out), cfg, keep_self=keep_self)
        c2 = swg.FilterConfig(**vars(cfg))
        c2.keep_self = keep_self
        oracle_lib.filter_paf(c2, str(src), str(ref))
        assert out.read_bytes() == ref.read_bytes(), keep_self
        outs[keep_self] = out.read_bytes()
    assert outs[False] != outs[True]
    path = swg.apply_paf_filter(str(src), cfg)
    try:
        assert path.endswith(".filtered.paf") and open(path, "rb").read() == outs[False]
    finally:
        os.unlink(path)
    aln = tmp_path / "x.1aln"
    aln.write_bytes(b"1 3 aln\n2 3 seq\n")
    with pytest.raises(swg.SwgError) as e:
        swg.filter_file(str(aln), str(tmp_path / "o.1aln"), cfg)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(swg.SwgError):
        swg.filter_file(str(tmp_path / "missing.paf"), str(tmp_path / "o.paf"), cfg)
    # .1aln in, PAF out, through an ALNtoPAF stand-in that prints `src` (the converter route of src/main.rs:737-770):
    # the same bytes as filtering the PAF itself; without a converter, or for .1aln output, UNSUPPORTED
    monkeypatch.delenv("SWG_ALNTOPAF", raising=False)
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    with pytest.raises(swg.SwgError) as e:
        swg.filter_file(str(aln), str(tmp_path / "o_aln.paf"), cfg)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    exe = tmp_path / "ALNtoPAF"
    exe.write_text(f'#!/bin/sh\n[ "$1" = "-x" ] || exit 4\n/bin/cat "{src}"\n')
    exe.chmod(0o755)
    monkeypatch.setenv("SWG_ALNTOPAF", str(exe))
    swg.filter_file(str(aln), str(tmp_path / "o_aln.paf"), cfg)
    assert (tmp_path / "o_aln.paf").read_bytes() == outs[False]
    assert not (tmp_path / "o_aln.paf.swg-aln.paf").exists()
    with pytest.raises(swg.SwgError) as e:
        swg.filter_file(str(aln), str(tmp_path / "o2.1aln"), cfg)
    assert e.value.code == _lib.ERR_UNSUPPORTED


def test_filter_paf_device_no_records_and_unwritable(ctx, tmp_path):
    (tmp_path / "junk.paf").write_bytes(b"\nnot a paf line\n")
    f = swg.PafFilter(swg.FilterConfig())
    f._ctx = ctx
    out = tmp_path / "o.paf"
    f.filter_paf(str(tmp_path / "junk.paf"), str(out))
    assert out.read_bytes() == b""
    t = synth.yeast_like(500, seed=1)
    synth.write_paf(t, str(tmp_path / "s.paf"))
    with pytest.raises(swg.SwgError):
        f.filter_paf(str(tmp_path / "s.paf"), str(tmp_path / "no_such_dir" / "o.paf"))
