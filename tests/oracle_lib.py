"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference filter.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
It does NOT import the product package: the two POD layouts it needs (swg_config, swg_stats of
include/sweepga_b200.h) are declared here, so that a process that only runs the oracle (bench.py --impl reference)
never maps libsweepga_b200.so.  tests/test_host.py checks the two declarations against the product binding's.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
_lib = C.CDLL(_PATH)

NO_LIMIT = (1 << 64) - 1


class swg_config(C.Structure):  # include/sweepga_b200.h: swg_config
    _fields_ = [
        ("min_block_length", C.c_uint64), ("mapping_max_per_query", C.c_uint64), ("mapping_max_per_target", C.c_uint64),
        ("scaffold_max_per_query", C.c_uint64), ("scaffold_max_per_target", C.c_uint64), ("scaffold_gap", C.c_uint64),
        ("min_scaffold_length", C.c_uint64), ("scaffold_max_deviation", C.c_uint64),
        ("overlap_threshold", C.c_double), ("scaffold_overlap_threshold", C.c_double),
        ("min_identity", C.c_double), ("min_scaffold_identity", C.c_double),
        ("mapping_filter_mode", C.c_uint8), ("scaffold_filter_mode", C.c_uint8), ("scoring_function", C.c_uint8),
        ("keep_self", C.c_uint8), ("scaffolds_only", C.c_uint8), ("reserved", C.c_uint8 * 3),
    ]


class swg_stats(C.Structure):  # include/sweepga_b200.h: swg_stats
    _fields_ = [
        ("n_input", C.c_uint64), ("n_stage1", C.c_uint64), ("n_after_sweep", C.c_uint64), ("n_chains", C.c_uint64),
        ("n_chains_after_mass", C.c_uint64), ("n_chains_kept", C.c_uint64), ("n_anchors", C.c_uint64),
        ("n_rescued", C.c_uint64), ("n_kept", C.c_uint64), ("score_near_ties", C.c_uint64), ("gpu_launches", C.c_uint64),
        ("ms_h2d", C.c_double), ("ms_device", C.c_double), ("ms_d2h", C.c_double),
        ("ms_sort_passes", C.c_double), ("n_sort_passes", C.c_uint64), ("n_sort_pairs", C.c_uint64),
        ("ms_tokenize", C.c_double), ("ms_write", C.c_double), ("exact_rerank", C.c_uint64), ("sort_bytes_per_pair", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("n_dirty_groups", C.c_uint64),
        ("ms_prefilter", C.c_double), ("prefilter_bytes_per_record", C.c_uint64), ("n_unsorted_groups", C.c_uint64),
    ]


class Config:
    """Stand-alone config for oracle-only processes: CLI defaults (src/cli.rs:204-276) unless overridden by keyword
    (field names of swg_config; None = no limit)."""

    def __init__(self, **kw):
        c = swg_config()
        c.mapping_max_per_query = c.mapping_max_per_target = c.scaffold_max_per_query = c.scaffold_max_per_target = NO_LIMIT
        c.scaffold_gap, c.min_scaffold_length, c.scaffold_max_deviation = 50_000, 10_000, 0
        c.overlap_threshold, c.scaffold_overlap_threshold = 0.95, 0.5
        c.mapping_filter_mode = c.scaffold_filter_mode = 2
        c.scoring_function = 3
        for k, v in kw.items():
            setattr(c, k, NO_LIMIT if v is None else v)
        self._c = c

    def to_c(self):
        return self._c


u8p, u32p, u64p, f64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_double)
_lib.orc_apply_filters.restype = C.c_int
_lib.orc_apply_filters.argtypes = [C.c_void_p, C.c_uint64, u32p, u32p, u64p, u64p, u64p, u64p, u64p, u64p, f64p, u8p,
                                   u32p, u32p, u8p, u32p, C.c_void_p, u32p, u32p]
_lib.orc_plane_sweep.restype = C.c_int
_lib.orc_plane_sweep.argtypes = [C.c_int, C.c_uint64, u64p, u64p, u64p, u64p, f64p, C.c_uint64, C.c_uint64, C.c_double, C.c_int, u8p]
_lib.orc_score.restype = C.c_double
_lib.orc_score.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_int]
_lib.orc_plane_sweep_core.restype = C.c_int
_lib.orc_plane_sweep_core.argtypes = [C.c_uint64, u32p, u32p, f64p, C.c_uint64, C.c_double, u64p]
_lib.orc_paf_parse.restype = C.c_void_p
_lib.orc_paf_parse.argtypes = [C.c_char_p]
_lib.orc_paf_free.argtypes = [C.c_void_p]
for f, r in (("orc_paf_n", C.c_uint64), ("orc_paf_n_lines", C.c_uint64), ("orc_paf_n_seq", C.c_uint32)):
    getattr(_lib, f).restype = r
    getattr(_lib, f).argtypes = [C.c_void_p]
_lib.orc_paf_seq_name.restype = C.c_char_p
_lib.orc_paf_seq_name.argtypes = [C.c_void_p, C.c_uint32]
_lib.orc_paf_u64.restype = u64p
_lib.orc_paf_u64.argtypes = [C.c_void_p, C.c_int]
_lib.orc_paf_u32.restype = u32p
_lib.orc_paf_u32.argtypes = [C.c_void_p, C.c_int]
_lib.orc_paf_identity.restype = f64p
_lib.orc_paf_identity.argtypes = [C.c_void_p]
_lib.orc_paf_strand.restype = u8p
_lib.orc_paf_strand.argtypes = [C.c_void_p]
_lib.orc_paf_write.restype = C.c_int
_lib.orc_paf_write.argtypes = [C.c_void_p, C.c_char_p, u8p, u32p]
_lib.orc_filter_paf.restype = C.c_int
_lib.orc_filter_paf.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p]

USIZE_MAX = (1 << 64) - 1


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def apply_filters(cfg, table, with_chain_keys=False):
    """Oracle apply_filters on a sweepga_b200.MappingTable -> (status u8[n], chain_id u32[n], stats[, keyA, keyB])."""
    n = table.n
    c = cfg.to_c()
    a64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    qs, qe, ts, te = a64(table.query_start), a64(table.query_end), a64(table.target_start), a64(table.target_end)
    bl, mt = a64(table.block_length), a64(table.matches)
    status, chain = np.zeros(n, np.uint8), np.zeros(n, np.uint32)
    st = swg_stats()
    kA = np.zeros(n + 1, np.uint32) if with_chain_keys else None
    kB = np.zeros(n + 1, np.uint32) if with_chain_keys else None
    _lib.orc_apply_filters(C.byref(c), n, _p(table.query_id, C.c_uint32), _p(table.target_id, C.c_uint32), _p(qs, C.c_uint64),
                           _p(qe, C.c_uint64), _p(ts, C.c_uint64), _p(te, C.c_uint64), _p(bl, C.c_uint64), _p(mt, C.c_uint64),
                           _p(table.identity, C.c_double), _p(table.strand, C.c_uint8), _p(table.seq_genome_id, C.c_uint32),
                           _p(table.seq_genome2_id, C.c_uint32), _p(status, C.c_uint8), _p(chain, C.c_uint32), C.byref(st),
                           _p(kA, C.c_uint32) if with_chain_keys else None, _p(kB, C.c_uint32) if with_chain_keys else None)
    if with_chain_keys:
        return status, chain, st, kA[: st.n_chains_kept], kB[: st.n_chains_kept]
    return status, chain, st


def plane_sweep(axis, mappings, n_keep, thr, scoring=3, n_keep2=None):
    """axis 'query' | 'target' | 'both'; mappings = [(qs, qe, ts, te, identity), ...] -> kept local indices."""
    ax = {"query": 0, "target": 1, "both": 2}[axis]
    cols = [np.ascontiguousarray([m[k] for m in mappings], dtype=np.uint64) for k in range(4)]
    idy = np.ascontiguousarray([m[4] for m in mappings], dtype=np.float64)
    keep = np.zeros(len(mappings), np.uint8)
    nk = USIZE_MAX if n_keep is None else n_keep
    nk2 = USIZE_MAX if n_keep2 is None else n_keep2
    _lib.orc_plane_sweep(ax, len(mappings), *[_p(c, C.c_uint64) for c in cols], _p(idy, C.c_double), nk, nk2, thr, scoring, _p(keep, C.c_uint8))
    return [int(i) for i in np.nonzero(keep)[0]]


def score(qs, qe, identity, scoring=3):
    return _lib.orc_score(qs, qe, identity, scoring)


_lib.orc_score_column.restype = None
_lib.orc_score_column.argtypes = [C.c_uint64, u32p, u32p, f64p, C.c_int, f64p]
_lib.orc_chain_identity.restype = None
_lib.orc_chain_identity.argtypes = [C.c_uint64, u64p, u64p, u64p, f64p]


def score_column(qs, qe, identity, scoring=3):
    """score_with_function over columns (host libm)."""
    qs, qe = np.ascontiguousarray(qs, np.uint32), np.ascontiguousarray(qe, np.uint32)
    identity = np.ascontiguousarray(identity, np.float64)
    out = np.empty(len(qs), np.float64)
    _lib.orc_score_column(len(qs), _p(qs, C.c_uint32), _p(qe, C.c_uint32), _p(identity, C.c_double), scoring, _p(out, C.c_double))
    return out


def chain_identity(total_length, sum_block, sum_matches):
    """weighted_identity (paf_filter.rs:896-913) over columns (host libm)."""
    a = [np.ascontiguousarray(x, np.uint64) for x in (total_length, sum_block, sum_matches)]
    out = np.empty(len(a[0]), np.float64)
    _lib.orc_chain_identity(len(a[0]), *[_p(x, C.c_uint64) for x in a], _p(out, C.c_double))
    return out


def plane_sweep_core(intervals, max_keep, thr):
    """intervals = [(begin, end, score), ...] -> kept indices in the reference's output order"""
    b = np.ascontiguousarray([i[0] for i in intervals], dtype=np.uint32)
    e = np.ascontiguousarray([i[1] for i in intervals], dtype=np.uint32)
    s = np.ascontiguousarray([i[2] for i in intervals], dtype=np.float64)
    out = np.zeros(max(len(intervals), 1), np.uint64)
    k = _lib.orc_plane_sweep_core(len(intervals), _p(b, C.c_uint32), _p(e, C.c_uint32), _p(s, C.c_double),
                                  USIZE_MAX if max_keep is None else max_keep, thr, _p(out, C.c_uint64))
    return [int(x) for x in out[:k]]


def parse_paf(path):
    """Oracle extract_metadata -> sweepga_b200.MappingTable with u64 coordinates kept in .coords64"""
    from sweepga_b200 import MappingTable
    h = _lib.orc_paf_parse(os.fsencode(path))
    if not h:
        raise IOError(path)
    try:
        n, ns = _lib.orc_paf_n(h), _lib.orc_paf_n_seq(h)
        g64 = lambda w: np.ctypeslib.as_array(_lib.orc_paf_u64(h, w), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        g32 = lambda w, cnt: np.ctypeslib.as_array(_lib.orc_paf_u32(h, w), shape=(cnt,)).copy() if cnt else np.zeros(0, np.uint32)
        ident = np.ctypeslib.as_array(_lib.orc_paf_identity(h), shape=(n,)).copy() if n else np.zeros(0)
        strand = np.ctypeslib.as_array(_lib.orc_paf_strand(h), shape=(n,)).copy() if n else np.zeros(0, np.uint8)
        cols = {k: g64(i) for i, k in enumerate(["rank", "qs", "qe", "ts", "te", "blen", "matches"])}
        t = MappingTable(g32(0, n), g32(1, n), cols["qs"], cols["qe"], cols["ts"], cols["te"], cols["blen"], cols["matches"], ident,
                         strand, g32(2, ns), g32(3, ns))
        t.names = [_lib.orc_paf_seq_name(h, i).decode() for i in range(ns)]
        t.rank = cols["rank"]
        t.n_lines = _lib.orc_paf_n_lines(h)
        return t
    finally:
        _lib.orc_paf_free(h)


_lib.orc_ani_stats.restype = C.c_int
_lib.orc_ani_stats.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, f64p, u64p]


def ani_stats(path, method, percentile=0.0, sort=1):
    """Oracle calculate_ani_stats (src/main.rs:334-688) -> (ani50, n_pairs); raises ValueError where the reference panics."""
    ani, npairs = C.c_double(), C.c_uint64()
    rc = _lib.orc_ani_stats(os.fsencode(path), method, percentile, sort, C.byref(ani), C.byref(npairs))
    if rc == -2:
        raise ValueError("NaN (the reference panics)")
    if rc != 0:
        raise IOError(path)
    return ani.value, int(npairs.value)


_lib.orc_tree_filter_paf.restype = C.c_int
_lib.orc_tree_filter_paf.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_double, u64p, u64p]
_lib.orc_siphash13.restype = C.c_uint64
_lib.orc_siphash13.argtypes = [C.c_char_p, C.c_uint64]


def tree_filter_paf(in_path, out_path, k_nearest, k_farthest=0, random_fraction=0.0):
    """Oracle apply_tree_filter_to_paf (src/tree_filter.rs:205-283) -> (lines kept, pairs selected)."""
    kept, sel = C.c_uint64(), C.c_uint64()
    rc = _lib.orc_tree_filter_paf(os.fsencode(in_path), os.fsencode(out_path), k_nearest, k_farthest, random_fraction, C.byref(kept), C.byref(sel))
    if rc == -2:
        raise ValueError("NaN (the reference panics)")
    if rc != 0:
        raise IOError(in_path)
    return int(kept.value), int(sel.value)


def siphash13(data: bytes) -> int:
    return int(_lib.orc_siphash13(data, len(data)))


def filter_paf(cfg, in_path, out_path):
    """Oracle PafFilter::filter_paf (parse + filter + tagged write) -> stats"""
    c = cfg.to_c()
    st = swg_stats()
    rc = _lib.orc_filter_paf(C.byref(c), os.fsencode(in_path), os.fsencode(out_path), C.byref(st))
    if rc != 0:
        raise IOError(in_path)
    return st
