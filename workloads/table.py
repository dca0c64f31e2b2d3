"""Column table of mappings (the SoA of include/sweepga_b200.h, as numpy arrays) — pure numpy.

`sweepga_b200.MappingTable` derives from `Table` and adds the ctypes view; the oracle binding
(tests/oracle_lib.py) and the reference arm of bench.py use `Table` directly.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np

U32_COLUMNS = ("query_id", "target_id", "query_start", "query_end", "target_start", "target_end", "block_length", "matches",
               "seq_genome_id", "seq_genome2_id")


@dataclass
class Table:
    query_id: np.ndarray
    target_id: np.ndarray
    query_start: np.ndarray
    query_end: np.ndarray
    target_start: np.ndarray
    target_end: np.ndarray
    block_length: np.ndarray
    matches: np.ndarray
    identity: np.ndarray
    strand: np.ndarray              # uint8: ord('+') forward, anything else reverse
    seq_genome_id: np.ndarray       # per sequence: id of P(name)   (src/paf_filter.rs:1022-1030)
    seq_genome2_id: np.ndarray      # per sequence: id of P2(name)  (src/plane_sweep_scaffold.rs:13-22)
    score: Optional[np.ndarray] = None
    names: Optional[list] = None
    rank: Optional[np.ndarray] = None  # PAF line number of each record (parse only)

    def __post_init__(self):
        for f in U32_COLUMNS:
            setattr(self, f, np.ascontiguousarray(getattr(self, f), dtype=np.uint32))
        self.identity = None if self.identity is None else np.ascontiguousarray(self.identity, dtype=np.float64)
        self.strand = np.ascontiguousarray(self.strand, dtype=np.uint8)
        if self.score is not None:
            self.score = np.ascontiguousarray(self.score, dtype=np.float64)

    @property
    def n(self):
        return int(self.query_id.shape[0])

    @property
    def n_seq(self):
        return int(self.seq_genome_id.shape[0])

    def take(self, idx):
        """Sub-table of the given record indices (sequence table shared)."""
        g = lambda a: a[idx]
        return type(self)(g(self.query_id), g(self.target_id), g(self.query_start), g(self.query_end), g(self.target_start),
                          g(self.target_end), g(self.block_length), g(self.matches), None if self.identity is None else g(self.identity),
                          g(self.strand), self.seq_genome_id, self.seq_genome2_id, None if self.score is None else g(self.score), self.names)

    @classmethod
    def from_names(cls, qnames, tnames, qs, qe, ts, te, blen, matches, identity, strand):
        """Intern names (first-appearance ids, one shared table) and derive P / P2 prefix ids."""
        ids, names = {}, []

        def iid(s):
            if s not in ids:
                ids[s] = len(names)
                names.append(s)
            return ids[s]
        q = np.empty(len(qnames), np.uint32)
        t = np.empty(len(qnames), np.uint32)
        for i, (a, b) in enumerate(zip(qnames, tnames)):
            q[i] = iid(a)
            t[i] = iid(b)
        P, P2 = prefix_ids(names)
        st = np.array([ord("+") if s == "+" else ord("-") for s in strand], np.uint8)
        return cls(q, t, qs, qe, ts, te, blen, matches, identity, st, P, P2, None, names)


def concat(tables):
    """Row-wise concatenation of tables that share one sequence table."""
    t0 = tables[0]
    cat = lambda f: np.concatenate([getattr(t, f) for t in tables])
    return type(t0)(cat("query_id"), cat("target_id"), cat("query_start"), cat("query_end"), cat("target_start"), cat("target_end"),
                    cat("block_length"), cat("matches"), None if t0.identity is None else cat("identity"), cat("strand"),
                    t0.seq_genome_id, t0.seq_genome2_id, None if t0.score is None else cat("score"), t0.names)


def prefix_P(name: str) -> str:
    """src/paf_filter.rs:1022-1030"""
    p = name.rfind("#")
    return name if p < 0 else name[: p + 1]


def prefix_P2(name: str) -> str:
    """src/plane_sweep_scaffold.rs:13-22"""
    parts = name.split("#")
    return f"{parts[0]}#{parts[1]}#" if len(parts) >= 2 else name


def prefix_ids(names):
    pid, p2id = {}, {}
    P = np.array([pid.setdefault(prefix_P(n), len(pid)) for n in names], np.uint32)
    P2 = np.array([p2id.setdefault(prefix_P2(n), len(p2id)) for n in names], np.uint32)
    return P, P2


def lpt_shards(table: Table, n_shards: int):
    """Size-balanced (longest-processing-time-first) assignment of genome-pair units (P(q), P(t)) to shards, in numpy —
    the same rule as swg_shard_plan (largest unit first, ties by first appearance, onto the least loaded shard, ties
    by lowest shard).  Returns (shard_of[n], shard_sizes[n_shards])."""
    import heapq
    # sequence classes closed under BOTH prefix rules (P and P2), like swg_shard_plan
    ns = table.n_seq
    parent = list(range(ns))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for col in (table.seq_genome_id, table.seq_genome2_id):
        first = {}
        for i, v in enumerate(col.tolist()):
            j = first.setdefault(v, i)
            if j != i:
                parent[find(i)] = find(j)
    P = np.array([find(i) for i in range(ns)], np.uint64)
    unit = (P[table.query_id] << np.uint64(32)) | P[table.target_id]
    uniq, first, inv, counts = np.unique(unit, return_index=True, return_inverse=True, return_counts=True)
    order = sorted(range(len(uniq)), key=lambda u: (-int(counts[u]), int(first[u])))
    heap = [(0, s) for s in range(n_shards)]
    shard_of_unit = np.zeros(len(uniq), np.uint32)
    sizes = np.zeros(n_shards, np.uint64)
    for u in order:
        load, s = heapq.heappop(heap)
        shard_of_unit[u] = s
        sizes[s] += np.uint64(counts[u])
        heapq.heappush(heap, (load + int(counts[u]), s))
    return shard_of_unit[inv].astype(np.uint32), sizes
