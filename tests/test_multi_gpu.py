"""Multi-GPU tests on real devices (need >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).

* two contexts on two devices in ONE process (the one-sweep kernels' shared-memory opt-in is a per-device attribute);
* the N > 1 scheme of bench.py on two NCCL ranks: ONE unit-structured table, each rank filters its own units on its GPU,
  (A, count) runs are exchanged, chain numbers renumbered on the device (swg_last_chain_units / swg_renumber_chains_device),
  2-bit status planes gathered — the merged result equals the single-GPU run of the whole table, chain numbers included.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")


@pytest.mark.gpu
@needs2
def test_two_contexts_two_devices_one_process():
    import oracle_lib
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    t = synth.yeast_like(20000, seed=3)
    cfg = swg.FilterConfig.from_cli(num_mappings="1:1", scaffold_filter="1:1")
    ref = oracle_lib.apply_filters(cfg, t)
    with swg.Context(0) as c0, swg.Context(1) as c1:
        for c in (c0, c1, c0, c1):
            status, chain, _ = c.filter(cfg, t)
            assert np.array_equal(status, ref[0]) and np.array_equal(chain, ref[1])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch
    import torch.distributed as dist
    import sweepga_b200 as swg
    from workloads import synth as wsynth
    from sweepga_b200.distributed import gather_runs, unit_offsets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n_total, seed, n_hap = 400_000, 4, 6
    pairs, quota = wsynth.pansn_unit_plan(n_total, n_hap)
    shard_of_unit, _ = swg.shard_plan_units(quota, world)
    mine = np.nonzero(shard_of_unit == rank)[0]
    wt, sizes = wsynth.pansn_units(mine, n_total, seed, n_hap)
    sub = swg.MappingTable(wt.query_id, wt.target_id, wt.query_start, wt.query_end, wt.target_start, wt.target_end, wt.block_length,
                           wt.matches, wt.identity, wt.strand, wt.seq_genome_id, wt.seq_genome2_id)
    all_sizes = torch.zeros(len(pairs), dtype=torch.int64, device=dev)
    all_sizes[torch.from_numpy(mine).to(dev)] = torch.from_numpy(sizes).to(dev)
    dist.all_reduce(all_sizes)
    goff = np.concatenate(([0], np.cumsum(all_sizes.cpu().numpy())))
    loff = np.concatenate(([0], np.cumsum(sizes)))
    cfg = swg.FilterConfig.from_cli(scaffold_dist="50k")
    n = sub.n
    with swg.Context(rank) as ctx:
        d_in, d_res = ctx.upload(sub)
        status_t = torch.zeros(n, dtype=torch.uint8, device=dev)
        chain_t = torch.zeros(n, dtype=torch.int32, device=dev)
        d_res.status = C.cast(status_t.data_ptr(), C.POINTER(C.c_uint8))
        d_res.chain_id = C.cast(chain_t.data_ptr(), C.POINTER(C.c_uint32))
        st = ctx.filter_device(cfg, d_in, d_res)
        a_loc, first_k = ctx.last_chain_units()
        cnt = np.diff(np.concatenate((first_k.astype(np.int64), [int(st.n_chains_kept) + 1])))
        u = np.searchsorted(loff, a_loc, side="right") - 1
        a_glob = goff[mine[u]] + (a_loc.astype(np.int64) - loff[u])
        runs = gather_runs(dist, a_glob, cnt, world, device=dev)
        delta = unit_offsets(runs)[rank]
        ctx.renumber_chains_device(n, chain_t.data_ptr(), first_k, delta)
        n_words = (n + 15) // 16
        packed = torch.zeros(n_words, dtype=torch.int32, device=dev)
        ctx.pack_status_device(n, status_t.data_ptr(), packed.data_ptr())
        torch.cuda.synchronize()
        w = packed.cpu().numpy().view(np.uint32)
        unpacked = ((w[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).reshape(-1)[:n].astype(np.uint8)
        assert np.array_equal(unpacked, status_t.cpu().numpy())
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), status=status_t.cpu().numpy(), chain=chain_t.cpu().numpy().view(np.uint32), units=mine)
    dist.destroy_process_group()


@pytest.mark.gpu
@needs2
def test_two_nccl_ranks_reproduce_the_single_gpu_run(tmp_path):
    import torch.multiprocessing as mp
    import sweepga_b200 as swg
    from workloads import synth as wsynth
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    pairs, quota = wsynth.pansn_unit_plan(400_000, 6)
    wt, sizes = wsynth.pansn_units(range(len(pairs)), 400_000, 4, 6)
    whole = swg.MappingTable(wt.query_id, wt.target_id, wt.query_start, wt.query_end, wt.target_start, wt.target_end, wt.block_length,
                             wt.matches, wt.identity, wt.strand, wt.seq_genome_id, wt.seq_genome2_id)
    with swg.Context(0) as ctx:
        ref_s, ref_c, _ = ctx.filter(swg.FilterConfig.from_cli(scaffold_dist="50k"), whole)
    goff = np.concatenate(([0], np.cumsum(sizes)))
    got_s, got_c = np.zeros(whole.n, np.uint8), np.zeros(whole.n, np.uint32)
    for r in range(2):
        d = np.load(tmp_path / f"r{r}.npz")
        idx = np.concatenate([np.arange(goff[u], goff[u + 1]) for u in d["units"]])
        got_s[idx], got_c[idx] = d["status"], d["chain"]
    assert np.array_equal(got_s, ref_s) and np.array_equal(got_c, ref_c)
    assert int((ref_c > 0).sum()) > 10000


@pytest.mark.gpu
@needs2
@pytest.mark.parametrize("flags", [{}, dict(num_mappings="1:1", scaffold_filter="1:1"), dict(scaffold_dist="50k")])
def test_swg_multi_filter_equals_single_gpu(flags):
    """Several GPUs behind ONE C-ABI call (swg_multi_filter): same status and chain numbers as one GPU."""
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    t = synth.pansn(300_000, seed=17, n_hap=8, with_names=True)
    cfg = swg.FilterConfig.from_cli(**flags)
    with swg.Context(0) as ctx:
        ref_s, ref_c, ref_st = ctx.filter(cfg, t)
    with swg.MultiContext([0, 1]) as mc:
        for _ in range(2):
            s, c, st = mc.filter(cfg, t)
            assert np.array_equal(s, ref_s) and np.array_equal(c, ref_c)
            assert st.n_kept == ref_st.n_kept and st.n_chains_kept == ref_st.n_chains_kept


@pytest.mark.gpu
def test_swg_multi_filter_one_device():
    """The same entry point with a single device (runs on the 1-GPU test box): the split / merge path with one shard."""
    import oracle_lib
    import sweepga_b200 as swg
    from sweepga_b200 import synth
    t = synth.yeast_like(20000, seed=5)
    cfg = swg.FilterConfig.from_cli(scaffold_dist="20k")
    ref = oracle_lib.apply_filters(cfg, t)
    with swg.MultiContext([0]) as mc:
        s, c, _ = mc.filter(cfg, t)
    assert np.array_equal(s, ref[0]) and np.array_equal(c, ref[1])
