#!/usr/bin/env python
"""bench.py — Mmappings/s of the mapping filter (BASELINE.json metric) on N B200s of one node.

A "step" is one pass of the filter (stage-1 retain -> plane sweep -> chaining -> scaffold filter/sweep ->
anchors/inversions [-> rescue]) over one batch of synthetic mappings.
  N = 1 : configs[2]  "synthetic PanSN PAF, 90 haplotypes x 24 chromosomes, 20M mappings, default pipeline"
  N > 1 : configs[3]  the same generator sharded by genome pair, 25M mappings per GPU (200M at 8), with
          --scaffold-dist 100k rescue; weak scaling, no data-path collective except the keep-bitmap gather.
`value`  : device-timed, inputs resident in HBM (CUDA events on the library's stream, max over ranks).
`e2e`    : the same through swg_filter with pinned HOST buffers, H2D and D2H inside the timed region.
`--impl reference` : the CPU oracle (the Rust reference cannot be built here) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][2]) if self.rows[0][2].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def make_workload(n_gpus, rank, records):
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    if n_gpus == 1:
        n = records or 20_000_000
        table = synth.pansn(n, seed=3)
        cfg = swg.FilterConfig()
        name = f"configs[2]: synthetic PanSN, 90 haplotypes x 24 chromosomes, {n} mappings, default pipeline"
        b_alg = 300
    else:
        n = records or 25_000_000
        table = synth.pansn(n, seed=4 + 1000 * rank)  # each rank owns whole genome pairs of the 200M-mapping job
        cfg = swg.FilterConfig.from_cli(scaffold_dist="100k")
        name = (f"configs[3]: synthetic PanSN sharded by genome pair, {n} mappings per GPU ({n * n_gpus} total), "
                "--scaffold-dist 100k")
        b_alg = 520
    return table, cfg, name, b_alg


def pinned_copy(table):
    """Column copies in page-locked memory (torch is only the allocator)."""
    import torch
    import sweepga_b200 as swg
    def pin(a):
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8, pin_memory=True)
        v = t.numpy()[: a.nbytes].view(a.dtype)
        v[...] = a
        return v, t
    keep = []
    cols = {}
    for f in ("query_id", "target_id", "query_start", "query_end", "target_start", "target_end", "block_length", "matches", "identity",
              "strand", "seq_genome_id", "seq_genome2_id"):
        v, t = pin(getattr(table, f))
        cols[f] = v
        keep.append(t)
    t2 = swg.MappingTable(**cols)
    t2._pins = keep
    st = torch.empty(max(table.n, 1), dtype=torch.uint8, pin_memory=True)
    ch = torch.empty(max(table.n, 1), dtype=torch.int32, pin_memory=True)
    t2._out = (st.numpy()[: table.n], ch.numpy()[: table.n].view(np.uint32), st, ch)
    return t2


def oracle_mt(cfg, table, threads):
    """The CPU restatement on `threads` host threads: genome-pair units (independent in the reference's algorithm)
    are size-balanced over threads; each thread runs the single-threaded filter on its units (ctypes releases the GIL)."""
    import oracle_lib
    import sweepga_b200 as swg
    from concurrent.futures import ThreadPoolExecutor
    if threads <= 1:
        t0 = time.perf_counter()
        oracle_lib.apply_filters(cfg, table)
        return time.perf_counter() - t0
    shard_of, _ = swg.shard_plan(table, threads)
    parts = [table.take(np.nonzero(shard_of == s)[0]) for s in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda p: oracle_lib.apply_filters(cfg, p), parts))
    return time.perf_counter() - t0


def paf_file_leg(ctx, cfg, n_lines):
    """Extra, informational: PafFilter::filter_paf file to file (text tokenised and tagged output assembled on the GPU,
    swg_filter_paf) and the same call with the host front end, on a synthetic PAF with cg:Z: tags.  Never fails the bench."""
    import shutil
    import tempfile
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    d = tempfile.mkdtemp(prefix="swg_bench_")
    try:
        t = synth.pansn(n_lines, seed=3, n_hap=40, with_names=True)
        src, out = os.path.join(d, "in.paf"), os.path.join(d, "out.paf")
        synth.write_paf_fast(t, src)
        f = swg.PafFilter(cfg)
        f._ctx = ctx
        res = {}
        for key, host in (("device_front_end", False), ("host_front_end", True)):
            f.filter_paf(src, out, host_frontend=host)  # warm-up: page cache, arenas, pinned pieces
            os.unlink(out)                               # a fresh output file (ext4 flushes a re-written truncated file at close)
            t0 = time.time()
            st = f.filter_paf(src, out, host_frontend=host)
            dt = time.time() - t0
            res[key] = {"wall_s": dt, "Mlines_per_s": t.n / dt / 1e6, "ms_upload": st.ms_h2d, "ms_tokenize": st.ms_tokenize,
                        "ms_filter": st.ms_device, "ms_write": st.ms_write, "gpu_launches": int(st.gpu_launches), "kept": int(st.n_kept)}
        res["lines"] = int(t.n)
        res["input_bytes"] = os.path.getsize(src)
        return res
    except Exception as e:  # informational leg only
        return {"error": repr(e)[:200]}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def skew_leg(ctx, cfg, n_pile):
    """Extra, informational: configs[4] (one centromeric pile + 100 k tiny groups) at --skew-pile mappings; the two strand
    groups of the pile go through the fixed-point chaining.  Host buffers, wall clock.  Never fails the bench."""
    try:
        from sweepga_b200 import synth
        t = synth.skew(n_pile=n_pile, n_tiny_groups=100_000, seed=5)
        ctx.filter(cfg, t)
        t0 = time.time()
        _, _, st = ctx.filter(cfg, t)
        dt = time.time() - t0
        return {"workload": f"configs[4] at reduced scale: {n_pile} mappings on one chromosome pair (95 % inside 6 Mbp) + 100000 tiny groups",
                "records": int(t.n), "wall_s": dt, "ms_device": float(st.ms_device), "Mmappings_per_s": t.n / dt / 1e6,
                "gpu_launches": int(st.gpu_launches), "kept": int(st.n_kept), "chains": int(st.n_chains_kept)}
    except Exception as e:  # informational leg only
        return {"error": repr(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=0, help="override the per-GPU record count (debug only)")
    ap.add_argument("--paf-lines", type=int, default=2_000_000,
                    help="N = 1 only: also time swg_filter_paf file to file on a synthetic PAF of this many lines (0 = skip)")
    ap.add_argument("--skew-pile", type=int, default=5_000_000,
                    help="size of the configs[4] pile of the informational skew leg (0 = skip)")
    ap.add_argument("--cpu-sample", type=int, default=20_000_000,
                    help="records of the workload the CPU oracle is timed on (~20 core-seconds at the default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus

    import __graft_entry__
    if not os.path.exists(__graft_entry__.LIB) or not os.path.exists(__graft_entry__.ORACLE):
        __graft_entry__.build()
    import sweepga_b200 as swg

    if args.impl == "reference":
        if rank != 0:
            return
        table, cfg, name, _ = make_workload(n_gpus, 0, args.records)
        sample = table.take(np.arange(min(table.n, args.cpu_sample)))
        cores = os.cpu_count() or 1
        times = []
        for i in range(args.warmup + args.steps):
            dt = oracle_mt(cfg, sample, cores)
            if i >= args.warmup:
                times.append(dt)
        per = sum(times) / len(times)
        v = sample.n / per / 1e6
        sample_desc = f"first {sample.n} records (whole genome pairs) of the workload, filter only, records in memory"
        print(json.dumps({
            "impl": "reference", "metric": "Mmappings/s filtered (sweep+scaffold)", "value": v, "unit": "Mmappings/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32/u64/f64", "data": "synthetic", "config": {"workload": name},
            "cpu_baseline": {"value": v, "unit": "Mmappings/s", "cores": cores, "kind": "port", "sample": sample_desc,
                             "note": "C++ oracle (statement-level restatement of the Rust reference, which cannot be built here: no cargo); "
                                     "the reference filter itself is single-threaded, this arm additionally spreads genome pairs over all host threads"},
            "e2e": {"value": v, "unit": "Mmappings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    table, cfg, name, b_alg = make_workload(n_gpus, rank, args.records)
    n = table.n
    ctx = swg.Context(local_rank)
    d_in, d_res = ctx.upload(table)

    # keep-bitmap gather (the only inter-GPU traffic on the path): status bytes of every shard to every rank
    gather_out = gather_in = None
    if world > 1:
        import ctypes as C
        gather_in = torch.empty(n, dtype=torch.uint8, device=dev)
        gather_out = torch.empty(n * world, dtype=torch.uint8, device=dev)
        d_res.status = C.cast(gather_in.data_ptr(), C.POINTER(C.c_uint8))
        chain_t = torch.empty(n, dtype=torch.int32, device=dev)
        d_res.chain_id = C.cast(chain_t.data_ptr(), C.POINTER(C.c_uint32))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        st = ctx.filter_device(cfg, d_in, d_res)
        ms = st.ms_device
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(gather_out, gather_in)
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms, st

    for _ in range(args.warmup):
        step_device()
    barrier()
    dev_ms, sort_ms, sort_passes, launches = 0.0, 0.0, 0, 0
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ms, st = step_device()
            dev_ms += ms
            sort_ms += st.ms_sort_passes
            sort_passes += st.n_sort_passes
            launches += st.gpu_launches
        barrier()
        wall_dev = time.perf_counter() - t0
    stats = st
    status_dev, chain_dev = ctx.download(n, d_res)

    # end to end: pinned host buffers in, host result out, copies inside the timed region
    pt = pinned_copy(table)
    out_s, out_c = pt._out[0], pt._out[1]
    for _ in range(min(args.warmup, 2)):
        ctx.filter(cfg, pt, out_s, out_c)
    barrier()
    with ClockSampler(local_rank) as clk2:  # the e2e loop is the longer timed region: more clock samples under load
        e2e_ms, t0 = 0.0, time.perf_counter()
        for _ in range(args.steps):
            _, _, st2 = ctx.filter(cfg, pt, out_s, out_c)
            e2e_ms += st2.ms_h2d + st2.ms_device + st2.ms_d2h
        barrier()
        wall_e2e = time.perf_counter() - t0
    clk.rows += clk2.rows
    assert np.array_equal(out_s, status_dev) and np.array_equal(out_c, chain_dev), "e2e and device-resident results differ"
    h2d = n * (8 * 4 + 8 + 1) + table.n_seq * 8
    d2h = n * 5

    t_dev, t_e2e = dev_ms / 1e3, max(e2e_ms / 1e3, 0.0)
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, wall_dev, wall_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, wall_dev, wall_e2e = tt.tolist()
        nn = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(nn)
        n_total = int(nn.item())
    else:
        n_total = n

    if rank == 0:
        peak, peak_src = peaks()
        value = n_total * args.steps / t_dev / 1e6
        e2e_v = n_total * args.steps / t_e2e / 1e6
        pass_ms = sort_ms / max(sort_passes, 1)
        achieved = stats.n_sort_pairs * 24 / (pass_ms / 1e3) / 1e9 if pass_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "Mmappings/s filtered (sweep+scaffold)", "value": value, "unit": "Mmappings/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64/f64", "data": "synthetic",
            "config": {"workload": name, "records_per_gpu": n, "l2": "inputs (41 B/record) are larger than the 126 MB L2",
                       "timing": "sum of per-step CUDA-event times on the library stream (+ NCCL gather events), max over ranks",
                       "wall_ms_per_step": wall_dev * 1e3 / args.steps,
                       "pipeline_algorithmic_bytes_per_mapping": b_alg,
                       "pipeline_fraction_of_hbm_roofline": (n * b_alg / (t_dev / args.steps)) / 1e9 / peak,
                       "stats": {k: int(getattr(stats, k)) for k in ("n_stage1", "n_after_sweep", "n_chains", "n_chains_after_mass",
                                                                     "n_chains_kept", "n_anchors", "n_rescued", "n_kept")}},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_v, "unit": "Mmappings/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e * 1e3 / args.steps, "wall_ms_per_step": wall_e2e * 1e3 / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "rs_onesweep_kernel (record sort pass, 12 B read + 12 B written per pair)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "launch_ms": pass_ms, "pairs_per_launch": int(stats.n_sort_pairs)},
        }
        if n_gpus == 1 and args.paf_lines > 0:
            line["paf_e2e"] = paf_file_leg(ctx, cfg, args.paf_lines)
        if n_gpus == 1 and args.skew_pile > 0:
            line["skew"] = skew_leg(ctx, cfg, args.skew_pile)
        if n_gpus == 1:
            sample = table.take(np.arange(min(n, args.cpu_sample)))
            cores = os.cpu_count() or 1
            dt = oracle_mt(cfg, sample, cores)
            line["cpu_baseline"] = {"value": sample.n / dt / 1e6, "unit": "Mmappings/s", "cores": cores, "kind": "port",
                                    "sample": f"first {sample.n} records (whole genome pairs) of the workload, filter only, {dt:.1f} s"}
        print(json.dumps(line))
    ctx.release(d_in, d_res) if world == 1 else None
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
