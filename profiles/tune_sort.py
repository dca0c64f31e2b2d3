#!/usr/bin/env python
"""Tuning aid: build libsweepga_b200 variants with different one-sweep tile shapes and time one pass each
(runs on the GPU box; nvcc is in the image).  Usage: python profiles/tune_sort.py 512x12 384x16 ..."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
for spec in sys.argv[1:]:
    parts = spec.split("x")
    th, it = parts[0], parts[1]
    lb = parts[2] if len(parts) > 2 else "8"
    mb = parts[3] if len(parts) > 3 else "2"
    extra = [f"-D{x}" for x in parts[4:]]
    out = f"/tmp/libswg_{spec}.so"
    cmd = ["nvcc"] + ge.NVCC_FLAGS + [f"-DSWG_RS_THREADS={th}", f"-DSWG_RS_ITEMS={it}", f"-DSWG_RS_LOOKBACK={lb}", f"-DSWG_RS_MINBLOCKS={mb}", *extra, "-shared", "-o", out] + \
          [os.path.join(ge.CSRC, s) for s in ge.SOURCES] + ["-lpthread", "-ldl", "-lrt", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(spec, "BUILD FAILED", r.stderr[-300:]); continue
    lib = C.CDLL(out)
    lib.swg_create.restype = C.c_void_p
    ctx = lib.swg_create(0)
    ms, ok = C.c_double(), C.c_int()
    for n, bits in ((20_000_000, 53),):
        rc = lib.swg__bench_sort(C.c_void_p(ctx), C.c_uint64(n), bits, 5, C.byref(ms), C.byref(ok))
        gbs = n * 24 / (ms.value / 1e3) / 1e9 if ms.value else 0
        print(f"{spec:12s} n={n} bits={bits} rc={rc} sorted={ok.value} ms/pass={ms.value:.4f}  {gbs:.0f} GB/s  frac={gbs/6543.1:.3f}", flush=True)
    lib.swg_destroy(C.c_void_p(ctx))
