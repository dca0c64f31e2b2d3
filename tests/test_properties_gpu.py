"""Size-independent properties of the CUDA path at sizes the oracle cannot reach in seconds
(BASELINE.json full sizes are exercised by bench.py; these run at a few million records)."""
import numpy as np
import pytest

import sweepga_b200 as swg
from sweepga_b200 import synth
from sweepga_b200.distributed import gpu_local_filter, merge_shards

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    return synth.pansn(3_000_000, seed=12, n_hap=30)


def test_determinism(ctx, big):
    cfg = swg.FilterConfig.from_cli(scaffold_dist="100k")
    a = ctx.filter(cfg, big)
    b = ctx.filter(cfg, big)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_device_resident_equals_host_path(ctx, big):
    cfg = swg.FilterConfig()
    s, c, _ = ctx.filter(cfg, big)
    dev, dres = ctx.upload(big)
    st = ctx.filter_device(cfg, dev, dres)
    s2, c2 = ctx.download(big.n, dres)
    ctx.release(dev, dres)
    assert np.array_equal(s, s2) and np.array_equal(c, c2)
    assert st.gpu_launches > 20


def test_idempotence_of_the_kept_set(ctx, big):
    """Filtering the kept records again keeps all of them, in the same chains (defaults: scaffold members only)."""
    cfg = swg.FilterConfig()
    s, c, _ = ctx.filter(cfg, big)
    kept = np.nonzero(s)[0]
    sub = big.take(kept)
    s2, c2, _ = ctx.filter(cfg, sub)
    assert (s2 != 0).all()
    # same partition into chains (numbering may shift because dropped records no longer define first appearances)
    _, inv1 = np.unique(c[kept], return_inverse=True)
    _, inv2 = np.unique(c2, return_inverse=True)
    pairs = np.unique(np.stack([inv1, inv2], axis=1), axis=0)
    assert len(pairs) == len(np.unique(inv1)) == len(np.unique(inv2))


def test_structural_invariants(ctx, big):
    cfg = swg.FilterConfig.from_cli(scaffold_dist="100k")
    s, c, st = ctx.filter(cfg, big)
    assert set(np.unique(s)) <= {0, 1, 2}
    assert ((s == 0) == (c == 0)).all()                      # kept <=> has a chain (scaffolding on)
    assert c.max() == st.n_chains_kept and st.n_kept == int((s != 0).sum())
    ids = np.unique(c[c > 0])
    assert np.array_equal(ids, np.arange(1, st.n_chains_kept + 1))   # chain_1..chain_K, no holes
    # every chain lives on one chromosome pair, and every rescued record's chain has an anchor on that pair
    key = big.query_id.astype(np.int64) * big.n_seq + big.target_id
    anchors = s == 1
    first = np.zeros(st.n_chains_kept + 1, np.int64)
    first[c[anchors]] = key[anchors]
    assert (first[c[anchors]] == key[anchors]).all()
    resc = s == 2
    assert (first[c[resc]] == key[resc]).all()
    # self mappings never survive without --self
    assert not (s[big.query_id == big.target_id] != 0).any()


@pytest.mark.parametrize("flags", [{}, dict(scaffold_dist="100k"), dict(num_mappings="1:1", scaffold_filter="1:1")])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_shard_invariance(ctx, flags, k):
    """Genome-pair shards filtered independently and merged == one shot (the multi-GPU contract), chain ids included."""
    table = synth.pansn(600_000, seed=13, n_hap=16)
    cfg = swg.FilterConfig.from_cli(**flags)
    ref_s, ref_c, _ = ctx.filter(cfg, table)
    shard_of, _ = swg.shard_plan(table, k)
    index = [np.nonzero(shard_of == s)[0] for s in range(k)]
    f = gpu_local_filter(ctx)
    res = [f(cfg, table.take(ix)) for ix in index]
    s, c = merge_shards(table.n, index, res)
    assert np.array_equal(s, ref_s)
    assert np.array_equal(c, ref_c)


def test_many_sequences_wide_keys(ctx):
    """Skewed table with thousands of contigs: exercises wide sequence-id fields of the sort keys."""
    import oracle_lib
    t = synth.skew(n_pile=3000, n_tiny_groups=20000, seed=8, window=2_000_000)
    for flags in ({}, dict(scaffold_mass="0", scaffold_dist="10k")):
        cfg = swg.FilterConfig.from_cli(**flags)
        s, c, _ = ctx.filter(cfg, t)
        os_, oc, _ = oracle_lib.apply_filters(cfg, t)
        assert np.array_equal(s, os_) and np.array_equal(c, oc)


def test_prefetch_gives_the_same_result(ctx, big):
    """swg_prefetch: a table uploaded ahead of its swg_filter call gives the same result; a filter call on other arrays ignores
    the prefetched table; at most two are outstanding; swg_prefetch_drop forgets them."""
    cfg = swg.FilterConfig.from_cli(scaffold_dist="100k")
    s0, c0, st0 = ctx.filter(cfg, big)
    ctx.prefetch(big)
    s1, c1, st1 = ctx.filter(cfg, big)            # consumes the prefetched copy
    assert np.array_equal(s0, s1) and np.array_equal(c0, c1) and st1.h2d_bytes == st0.h2d_bytes
    ctx.prefetch(big)
    ctx.prefetch(big)
    with pytest.raises(swg.SwgError):
        ctx.prefetch(big)                         # a third one
    other = big.take(np.arange(big.n // 2))
    other = swg.MappingTable(other.query_id, other.target_id, other.query_start, other.query_end, other.target_start, other.target_end,
                             other.block_length, other.matches, other.identity, other.strand, other.seq_genome_id, other.seq_genome2_id)
    so, co, _ = ctx.filter(cfg, other)            # different arrays: uploaded as usual, the prefetched tables stay
    s2, c2, _ = ctx.filter(cfg, big)
    s3, c3, _ = ctx.filter(cfg, big)
    assert np.array_equal(s0, s2) and np.array_equal(c0, c2) and np.array_equal(s0, s3) and np.array_equal(c0, c3)
    so2, co2, _ = ctx.filter(cfg, other)
    assert np.array_equal(so, so2) and np.array_equal(co, co2)
    ctx.prefetch(big)
    ctx.prefetch_drop()
    s4, c4, _ = ctx.filter(cfg, big)
    assert np.array_equal(s0, s4) and np.array_equal(c0, c4)


@pytest.mark.parametrize("n,bits", [(1, 8), (2, 64), (5375, 17), (5376, 33), (5377, 64), (100_000, 50), (2_097_151, 41), (2_097_152, 41), (3_000_000, 53)])
def test_radix_sort_tiles(ctx, n, bits):
    """The stable LSD sort of (key, payload) pairs on pseudo-random keys, checked on the device (sorted, stable, payload intact),
    at sizes around the tile boundaries (5376 pairs per CTA) and with 1 .. 8 passes."""
    import ctypes as C
    from sweepga_b200 import _lib
    fn = C.CDLL(_lib.LIB_PATH).swg__bench_sort
    fn.restype = C.c_int
    ms, ok = C.c_double(), C.c_int()
    rc = fn(C.c_void_p(ctx._h), C.c_uint64(n), C.c_int(bits), C.c_int(1), C.byref(ms), C.byref(ok))
    assert rc == 0 and ok.value == 1
