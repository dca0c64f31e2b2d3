#!/usr/bin/env python
"""File-level timing of the PAF front end (parse on all host threads -> GPU filter -> tagged write) against the
oracle's single-threaded restatement of filter_paf.  Usage: python profiles/bench_paf_frontend.py [n_records]"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sweepga_b200 as swg
from sweepga_b200 import synth
import oracle_lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
t = synth.pansn(n, seed=3, n_hap=40, with_names=True)
d = tempfile.mkdtemp()
src = os.path.join(d, "in.paf")
t0 = time.time()
# vectorised writer (synth.write_paf is a Python loop)
names = np.array(t.names)
cols = [names[t.query_id], (t.query_end + 1000).astype(str), t.query_start.astype(str), t.query_end.astype(str),
        np.where(t.strand == ord("+"), "+", "-"), names[t.target_id], (t.target_end + 1000).astype(str), t.target_start.astype(str),
        t.target_end.astype(str), t.matches.astype(str), t.block_length.astype(str), np.full(t.n, "60")]
lines = cols[0]
for c in cols[1:]:
    lines = np.char.add(np.char.add(lines, "\t"), c)
tags = np.char.add(np.char.add(np.char.add("\tcg:Z:", t.matches.astype(str)), "="), np.char.add((t.block_length - t.matches).astype(str), "X"))
lines = np.char.add(lines, tags)
with open(src, "w") as f:
    f.write("\n".join(lines.tolist()) + "\n")
print(f"wrote {t.n} lines, {os.path.getsize(src)/1e6:.0f} MB in {time.time()-t0:.1f} s", flush=True)
cfg = swg.FilterConfig()
ctx = swg.Context(0)
f = swg.PafFilter(cfg); f._ctx = ctx
out = os.path.join(d, "gpu.paf")
f.filter_paf(src, out)  # warm-up (page cache, arena, pinned pieces)
os.unlink(out)  # a fresh output file: re-writing a truncated file makes ext4 flush it to disk at close()
t0 = time.time(); st = f.filter_paf(src, out); t_gpu = time.time() - t0
outh = os.path.join(d, "host.paf")
f.filter_paf(src, outh, host_frontend=True)
os.unlink(outh)
t0 = time.time(); sth = f.filter_paf(src, outh, host_frontend=True); t_host = time.time() - t0
t0 = time.time(); tab = swg.parse_paf(src); t_parse = time.time() - t0
t0 = time.time(); tabd = swg.parse_paf(src, ctx); t_parse_dev = time.time() - t0
out2 = os.path.join(d, "orc.paf")
t_orc = float("nan")
if n <= 8_000_000:
    t0 = time.time(); oracle_lib.filter_paf(cfg, src, out2); t_orc = time.time() - t0
    same = open(out, "rb").read() == open(out2, "rb").read() == open(outh, "rb").read()
else:
    same = open(out, "rb").read() == open(outh, "rb").read()
mb = os.path.getsize(src) / 1e6
print(f"filter_paf, device front end: {t_gpu:.3f} s wall ({t.n/t_gpu/1e6:.2f} M lines/s, {mb/t_gpu/1e3:.2f} GB/s of text): upload {st.ms_h2d:.1f} ms, "
      f"tokenise {st.ms_tokenize:.1f} ms, filter {st.ms_device:.1f} ms, assemble+download+write {st.ms_write:.1f} ms, {st.gpu_launches} launches")
print(f"filter_paf, host front end:   {t_host:.3f} s wall ({t.n/t_host/1e6:.2f} M lines/s): filter {sth.ms_device:.1f} ms, h2d {sth.ms_h2d:.1f} ms")
print(f"parse alone incl. numpy copies: host {t_parse:.3f} s, device {t_parse_dev:.3f} s")
swg.ani_stats(ctx, src, "n100")
t0 = time.time(); ani = swg.ani_stats(ctx, src, "n100"); t_ani = time.time() - t0
t0 = time.time(); ani_all = swg.ani_stats(ctx, src, "all"); t_ani_all = time.time() - t0
if n <= 8_000_000:
    t0 = time.time(); o_ani = oracle_lib.ani_stats(src, 2, 100.0, 1); t_oani = time.time() - t0
    print(f"ANI pre-pass (n100-identity): device {t_ani:.3f} s, 'all' {t_ani_all:.3f} s | oracle 1 thread {t_oani:.3f} s | identical: {ani == o_ani} ({ani[0]!r}, {ani[1]} pairs)")
else:
    print(f"ANI pre-pass (n100-identity): device {t_ani:.3f} s, 'all' {t_ani_all:.3f} s ({ani[0]!r}, {ani[1]} pairs)")
print(f"oracle 1 thread: {t_orc:.3f} s ({t.n/t_orc/1e6:.2f} M lines/s) | identical output: {same}")
