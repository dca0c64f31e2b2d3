"""Multi-GPU driver logic: genome-pair shards, one process per GPU, no collective inside the filter.

The filter's groupings are all closed under the shard unit chosen by `shard_plan`, so every rank
filters its records independently; what crosses NVLink afterwards is the per-record result (one status byte and one
chain number) and, to reproduce the single-GPU `chain_N` numbering, two u32 order keys per kept chain
(`swg_last_chain_keys`).  `torch.distributed` is only the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

from .api import MappingTable, shard_plan


def merge_shards(n_total, shard_index, shard_results):
    """Pure merge step (no communication): shard_index[s] = original indices of shard s (ascending),
    shard_results[s] = (status, chain_local, keyA_local, keyB_local).  Returns global (status, chain_id)."""
    status = np.zeros(n_total, np.uint8)
    chain = np.zeros(n_total, np.uint32)
    keys = []
    for s, (idx, (st, ch, ka, kb)) in enumerate(zip(shard_index, shard_results)):
        k = len(ka)
        if k:
            a = idx[ka.astype(np.int64)].astype(np.int64)   # shard-local record index -> original index (monotone)
            b = idx[kb.astype(np.int64)].astype(np.int64)
            keys.append(np.stack([a, b, np.full(k, s, np.int64), np.arange(k, dtype=np.int64)], axis=1))
    if keys:
        allk = np.concatenate(keys, axis=0)
        order = np.lexsort((allk[:, 3], allk[:, 2], allk[:, 1], allk[:, 0]))  # by A, then B, then shard, then local k
        glob = np.empty(len(order), np.int64)
        glob[order] = np.arange(1, len(order) + 1)
        off = 0
        luts = []
        for s, (idx, (st, ch, ka, kb)) in enumerate(zip(shard_index, shard_results)):
            k = len(ka)
            lut = np.zeros(k + 1, np.uint32)
            lut[1:] = glob[off:off + k]
            off += k
            luts.append(lut)
    else:
        luts = [np.zeros(1, np.uint32) for _ in shard_index]
    for idx, (st, ch, ka, kb), lut in zip(shard_index, shard_results, luts):
        status[idx] = st
        chain[idx] = lut[ch]
    return status, chain


def _pad_gather(dist, t, world, group=None):
    """all_gather of 1-D tensors of different lengths (padded to the longest)."""
    import torch
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    m = max(max(sizes), 1)
    buf = torch.zeros(m, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def filter_sharded(local_filter, cfg, table: MappingTable, rank: int, world: int, device=None, group=None):
    """Every rank holds the same table (or at least the same sharding columns); rank r filters shard r with
    `local_filter(cfg, sub_table) -> (status, chain_local, keyA, keyB)` and all ranks return the global result.
    For GPUs pass `local_filter=gpu_local_filter(ctx)` and `device=torch.device('cuda', local_rank)`."""
    import torch
    import torch.distributed as dist
    shard_of, sizes = shard_plan(table, world)
    index = [np.nonzero(shard_of == s)[0] for s in range(world)]
    st, ch, ka, kb = local_filter(cfg, table.take(index[rank]))
    dev = device or torch.device("cpu")
    tt = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
    g_st = _pad_gather(dist, tt(st, np.uint8), world, group)
    g_ch = _pad_gather(dist, tt(ch, np.int64), world, group)
    g_ka = _pad_gather(dist, tt(ka, np.int64), world, group)
    g_kb = _pad_gather(dist, tt(kb, np.int64), world, group)
    res = [(g_st[s].cpu().numpy(), g_ch[s].cpu().numpy(), g_ka[s].cpu().numpy(), g_kb[s].cpu().numpy()) for s in range(world)]
    return merge_shards(table.n, index, res)


def gpu_local_filter(ctx):
    def f(cfg, sub):
        st, ch, _ = ctx.filter(cfg, sub)
        ka, kb = ctx.last_chain_keys()
        return st, ch, ka, kb
    return f


# ---- the cheaper merge: (A, count) per genome-pair unit instead of two keys per chain --------------------------------
def chain_runs(key_a):
    """Runs of equal A in a shard's chain order -> (A per run, first local chain number per run (1-based), run lengths).
    What swg_last_chain_units reports, from the per-chain keys (for fallbacks and tests)."""
    key_a = np.asarray(key_a)
    if len(key_a) == 0:
        z = np.zeros(0, np.int64)
        return z, z, z
    first = np.nonzero(np.concatenate(([True], key_a[1:] != key_a[:-1])))[0]
    return key_a[first].astype(np.int64), (first + 1).astype(np.int64), np.diff(np.concatenate((first, [len(key_a)]))).astype(np.int64)


def unit_offsets(all_runs):
    """all_runs[s] = (A_global per run, run length per run) of shard s, each in that shard's chain order.
    Returns delta[s][r] = (global number of the run's first chain) - (its local number): the kept chains of one
    genome-pair unit are consecutive in the local and in the merged numbering (order O3 sorts by A first), so a shard's
    chain k of run r becomes k + delta[s][r]."""
    a = np.concatenate([np.asarray(r[0], np.int64) for r in all_runs]) if all_runs else np.zeros(0, np.int64)
    n = np.concatenate([np.asarray(r[1], np.int64) for r in all_runs]) if all_runs else np.zeros(0, np.int64)
    order = np.argsort(a, kind="stable")
    excl = np.zeros(len(a), np.int64)
    excl[order] = np.cumsum(n[order]) - n[order]
    out, off = [], 0
    for r in all_runs:
        k = len(r[0])
        local_first = np.cumsum(np.asarray(r[1], np.int64)) - np.asarray(r[1], np.int64)  # 0-based
        out.append(excl[off:off + k] - local_first)
        off += k
    return out


def gather_runs(dist, a_global, counts, world, device=None, group=None):
    """all_gather of every rank's (A_global, count) runs -> list over ranks (tiny: one pair per genome-pair unit)."""
    import torch
    dev = device or torch.device("cpu")
    t = torch.from_numpy(np.stack([np.asarray(a_global, np.int64), np.asarray(counts, np.int64)], axis=1).reshape(-1)).to(dev)
    parts = _pad_gather(dist, t, world, group)
    return [(p.cpu().numpy().reshape(-1, 2)[:, 0], p.cpu().numpy().reshape(-1, 2)[:, 1]) for p in parts]
