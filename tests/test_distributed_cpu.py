"""N > 1 host logic on CPU: world_size-2 gloo processes shard by genome pair, filter their shard (the oracle stands in
for the GPU filter — allowed in tests), exchange results and must reproduce the single-shot result, chain numbers
included."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_local_filter(cfg, sub):
    import oracle_lib
    st, ch, _, ka, kb = oracle_lib.apply_filters(cfg, sub, with_chain_keys=True)
    return st, ch, ka, kb


def _worker(rank, world, port, flags, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    from sweepga_b200.distributed import filter_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = synth.yeast_like(12000, seed=21)
    cfg = swg.FilterConfig.from_cli(**flags)
    status, chain = filter_sharded(_oracle_local_filter, cfg, table, rank, world)
    np.save(os.path.join(out_dir, f"status_{rank}.npy"), status)
    np.save(os.path.join(out_dir, f"chain_{rank}.npy"), chain)
    dist.destroy_process_group()


@pytest.mark.parametrize("flags", [{}, dict(num_mappings="1:1", scaffold_filter="1:1"), dict(scaffold_dist="50k")])
def test_gloo_two_ranks_reproduce_single_shot(tmp_path, flags):
    import torch.multiprocessing as mp
    import oracle_lib
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, flags, str(tmp_path)), nprocs=2, join=True)
    table = synth.yeast_like(12000, seed=21)
    ref_s, ref_c, _ = oracle_lib.apply_filters(swg.FilterConfig.from_cli(**flags), table)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"status_{r}.npy"), ref_s)
        assert np.array_equal(np.load(tmp_path / f"chain_{r}.npy"), ref_c)


def test_merge_shards_many_ways():
    """merge_shards alone, 1..5 shards, including shards that end up empty."""
    import oracle_lib
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    from sweepga_b200.distributed import merge_shards
    table = synth.pansn(60000, seed=5, n_hap=6, with_names=True)
    for flags in ({}, dict(scaffold_filter="1:1", scaffold_mass="2k")):
        cfg = swg.FilterConfig.from_cli(**flags)
        ref_s, ref_c, _ = oracle_lib.apply_filters(cfg, table)
        for k in (1, 2, 3, 5):
            shard_of, _ = swg.shard_plan(table, k)
            index = [np.nonzero(shard_of == s)[0] for s in range(k)]
            res = [_oracle_local_filter(cfg, table.take(ix)) for ix in index]
            s, c = merge_shards(table.n, index, res)
            assert np.array_equal(s, ref_s) and np.array_equal(c, ref_c), (flags, k)


def test_unit_run_merge_equals_per_chain_merge():
    """unit_offsets (one (A, count) pair per genome-pair unit) numbers the chains exactly like merge_shards (two keys per
    chain) and like the single-shot run."""
    import oracle_lib
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    from sweepga_b200.distributed import chain_runs, unit_offsets
    table = synth.pansn(80000, seed=9, n_hap=6, with_names=True)
    for flags in ({}, dict(scaffold_filter="1:1", scaffold_mass="2k"), dict(scaffold_dist="30k")):
        cfg = swg.FilterConfig.from_cli(**flags)
        ref_s, ref_c, _ = oracle_lib.apply_filters(cfg, table)
        for k in (1, 2, 4, 7):
            shard_of, _ = swg.shard_plan(table, k)
            index = [np.nonzero(shard_of == s)[0] for s in range(k)]
            res = [_oracle_local_filter(cfg, table.take(ix)) for ix in index]
            runs = []
            for ix, (st, ch, ka, kb) in zip(index, res):
                a, first, cnt = chain_runs(ka)
                runs.append((ix[a] if len(a) else a, cnt))       # A: shard-local record index -> global index
            deltas = unit_offsets(runs)
            status, chain = np.zeros(table.n, np.uint8), np.zeros(table.n, np.uint32)
            for ix, (st, ch, ka, kb), d in zip(index, res, deltas):
                a, first, cnt = chain_runs(ka)
                lut = np.zeros(len(ka) + 1, np.int64)
                if len(ka):
                    run_of = np.repeat(np.arange(len(a)), cnt)
                    lut[1:] = np.arange(1, len(ka) + 1) + d[run_of]
                status[ix] = st
                chain[ix] = lut[ch]
            assert np.array_equal(status, ref_s) and np.array_equal(chain, ref_c), (flags, k)


def _unit_worker(rank, world, port, out_dir):
    """bench.py's N > 1 scheme end to end on CPU: ONE unit-structured table, each rank generates only its own units
    (swg_shard_plan_units on the planned sizes), filters them (oracle stands in), exchanges (A_global, count) runs."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib
    import sweepga_b200 as swg
    from workloads import synth as wsynth
    from sweepga_b200.distributed import chain_runs, gather_runs, unit_offsets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total, seed, n_hap = 60000, 4, 5
    pairs, quota = wsynth.pansn_unit_plan(n_total, n_hap)
    shard_of_unit, _ = swg.shard_plan_units(quota, world)
    mine = np.nonzero(shard_of_unit == rank)[0]
    sub, sizes = wsynth.pansn_units(mine, n_total, seed, n_hap)
    # global offsets of my units: every rank contributes the actual sizes of its units
    import torch
    all_sizes = torch.zeros(len(pairs), dtype=torch.int64)
    all_sizes[torch.from_numpy(mine)] = torch.from_numpy(sizes)
    dist.all_reduce(all_sizes)
    goff = np.concatenate(([0], np.cumsum(all_sizes.numpy())))
    loff = np.concatenate(([0], np.cumsum(sizes)))
    cfg = swg.FilterConfig.from_cli(scaffold_dist="50k")
    st, ch, _, ka, kb = oracle_lib.apply_filters(cfg, sub, with_chain_keys=True)
    a, first, cnt = chain_runs(ka)
    u = np.searchsorted(loff, a, side="right") - 1                 # local unit of record a
    a_glob = goff[mine[u]] + (a - loff[u])
    runs = gather_runs(dist, a_glob, cnt, world)
    delta = unit_offsets(runs)[rank]
    lut = np.zeros(len(ka) + 1, np.int64)
    if len(ka):
        lut[1:] = np.arange(1, len(ka) + 1) + delta[np.repeat(np.arange(len(a)), cnt)]
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), status=st, chain=lut[ch], units=mine, sizes=sizes)
    dist.destroy_process_group()


def test_gloo_unit_sharded_table_reproduces_single_shot(tmp_path):
    import torch.multiprocessing as mp
    import oracle_lib
    import sweepga_b200 as swg
    from workloads import synth as wsynth
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_unit_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    pairs, quota = wsynth.pansn_unit_plan(60000, 5)
    whole, sizes = wsynth.pansn_units(range(len(pairs)), 60000, 4, 5)
    ref_s, ref_c, _ = oracle_lib.apply_filters(swg.FilterConfig.from_cli(scaffold_dist="50k"), whole)
    goff = np.concatenate(([0], np.cumsum(sizes)))
    got_s, got_c = np.zeros(whole.n, np.uint8), np.zeros(whole.n, np.int64)
    for r in range(2):
        d = np.load(tmp_path / f"r{r}.npz")
        idx = np.concatenate([np.arange(goff[u], goff[u + 1]) for u in d["units"]])
        got_s[idx], got_c[idx] = d["status"], d["chain"]
    assert np.array_equal(got_s, ref_s) and np.array_equal(got_c, ref_c.astype(np.int64))
    assert int((ref_c > 0).sum()) > 1000
