#!/usr/bin/env python
"""Tuning aid: build libsweepga_b200 variants with different one-sweep tile shapes and time the record sort's passes
(runs on the GPU box; nvcc is in the image).  Usage: python profiles/tune_sort.py THREADSxITEMS[xLOOKBACK[xMINBLOCKS[xMINBLOCKS_PACKED[xDEFINE...]]]] ...
Prints, per variant: one pairs pass (24 B per pair) and one packed-word pass (16 B per word) on 20 M elements, 53 key bits."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
PEAK = 6543.1
for spec in sys.argv[1:]:
    parts = spec.split("x")
    th, it = parts[0], parts[1]
    lb = parts[2] if len(parts) > 2 else "8"
    mb = parts[3] if len(parts) > 3 else "2"
    mbp = parts[4] if len(parts) > 4 else "2"
    extra = [f"-D{x}" for x in parts[5:]]
    out = f"/tmp/libswg_{spec}.so"
    cmd = ["nvcc"] + ge.NVCC_FLAGS + [f"-DSWG_RS_THREADS={th}", f"-DSWG_RS_ITEMS={it}", f"-DSWG_RS_LOOKBACK={lb}", f"-DSWG_RS_MINBLOCKS={mb}",
                                      f"-DSWG_RS_MINBLOCKS_PACKED={mbp}", *extra, "-shared", "-o", out] + \
          [os.path.join(ge.CSRC, s) for s in ge.SOURCES] + ["-lpthread", "-ldl", "-lrt", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(spec, "BUILD FAILED", r.stderr[-300:]); continue
    lib = C.CDLL(out)
    lib.swg_create.restype = C.c_void_p
    ctx = lib.swg_create(0)
    ms, tot, ok = C.c_double(), C.c_double(), C.c_int()
    n, bits = 20_000_000, 53
    rc = lib.swg__bench_sort(C.c_void_p(ctx), C.c_uint64(n), bits, 5, C.byref(ms), C.byref(ok))
    gbs = n * 24 / (ms.value / 1e3) / 1e9 if ms.value else 0
    line = f"{spec:22s} pairs: rc={rc} ok={ok.value} {ms.value:.4f} ms/pass {gbs:5.0f} GB/s frac={gbs/PEAK:.3f}"
    rc = lib.swg__bench_sort_packed(C.c_void_p(ctx), C.c_uint64(n), bits, 5, C.byref(ms), C.byref(tot), C.byref(ok))
    gbs = n * 16 / (ms.value / 1e3) / 1e9 if ms.value else 0
    print(line + f" | packed: rc={rc} ok={ok.value} {ms.value:.4f} ms/pass {gbs:5.0f} GB/s frac={gbs/PEAK:.3f} whole sort {tot.value:.3f} ms", flush=True)
    lib.swg_destroy(C.c_void_p(ctx))
