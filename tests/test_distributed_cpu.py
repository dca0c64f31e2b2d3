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
