"""(pure numpy; no native code is loaded by importing this module)
Seeded synthetic mapping tables of the BASELINE.json shapes.

The real yeast PAF cannot be regenerated (FASTA blob missing, FastGA absent), so configs 1-2 use a
yeast-shaped stand-in with the genome/chromosome names and (SGDref) lengths of data/scerevisiae8.fa.gz.fai;
configs 3-5 are synthetic by definition.  Everything is vectorised numpy so that 20 M records generate in
seconds; records come out grouped by genome pair like aligner output.
"""
import numpy as np

from .table import Table, prefix_ids

YEAST_GENOMES = ["SGDref#1", "S288C#1", "DBVPG6044#1", "DBVPG6765#1", "SK1#1", "UWOPS034614#1", "Y12#1", "YPS128#1"]
YEAST_CHROMS = ["chrI", "chrII", "chrIII", "chrIV", "chrV", "chrVI", "chrVII", "chrVIII", "chrIX", "chrX", "chrXI", "chrXII",
                "chrXIII", "chrXIV", "chrXV", "chrXVI", "chrMT"]
YEAST_LENGTHS = [230218, 813184, 316620, 1531933, 576874, 270161, 1090940, 562643, 439888, 745751, 666816, 1078177, 924431,
                 784333, 1091291, 948066, 85779]
HUMAN_CHROMS = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]
HUMAN_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
                 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
                 46709983, 50818468, 156040895, 57227415]


def _segment_cumsum(x, seg_start_idx, seg_id):
    """exclusive cumulative sum of x restarted at each segment"""
    c = np.cumsum(x, dtype=np.int64) - x
    start = np.minimum(seg_start_idx, max(len(x) - 1, 0))  # empty trailing segments point one past the end (never used)
    return c - c[start][seg_id]


def pangenome(genomes, chroms, lengths, n_records, seed, block_mu, block_sigma, block_clip, self_genome=True,
              frac_offdiag=0.10, frac_inv=0.03, with_names=True, pairs=None):
    """All ordered genome pairs (or the given ones); collinear diagonals per same-chromosome pair + off-diagonal repeats."""
    rng = np.random.default_rng(seed)
    ng, nc = len(genomes), len(chroms)
    L = np.asarray(lengths, dtype=np.int64)
    if pairs is None:
        pairs = [(a, b) for a in range(ng) for b in range(ng) if self_genome or a != b]
    npairs = len(pairs)
    pa = np.array([p[0] for p in pairs], np.int64)
    pb = np.array([p[1] for p in pairs], np.int64)
    n_diag_total = int(round(n_records * (1.0 - frac_offdiag)))
    n_off = n_records - n_diag_total
    # diagonal records: (pair, chrom) diagonals with counts proportional to chromosome length
    w = np.tile(L / L.sum(), npairs) / npairs
    counts = rng.multinomial(n_diag_total, w)
    diag_of = np.repeat(np.arange(npairs * nc), counts)
    nd = diag_of.shape[0]
    seg_start = np.concatenate(([0], np.cumsum(counts)[:-1]))
    blen = np.clip(np.exp(rng.normal(block_mu, block_sigma, nd)), block_clip[0], block_clip[1]).astype(np.int64)
    gap = rng.exponential(2000.0, nd).astype(np.int64)
    ov = rng.random(nd) < 0.10
    gap[ov] = -rng.integers(0, 501, int(ov.sum()))
    step = np.maximum(blen + gap, 1)
    off = _segment_cumsum(step, seg_start, diag_of)
    extent = np.zeros(npairs * nc, np.int64)
    np.add.at(extent, diag_of, step)
    chrom_of_diag = np.tile(np.arange(nc), npairs)
    Ld = L[chrom_of_diag]
    room = np.maximum(Ld - extent - 2000, 1)
    start_q = (rng.random(npairs * nc) * room).astype(np.int64)
    start_t = np.clip(start_q + rng.integers(-1000, 1001, npairs * nc), 0, None)
    qs = start_q[diag_of] + off
    tlen = np.maximum(blen + rng.integers(-20, 21, nd), 50)
    ts = start_t[diag_of] + off + rng.integers(-30, 31, nd)
    ts = np.maximum(ts, 0)
    qe, te = qs + blen, ts + tlen
    ident = np.clip(1.0 - rng.exponential(0.008, nd), 0.70, 0.999999)
    strand = np.full(nd, ord("+"), np.uint8)
    strand[rng.random(nd) < frac_inv] = ord("-")
    pair_of = diag_of // nc
    qchr = chrom_of_diag[diag_of]
    tchr = qchr.copy()
    # off-diagonal / inter-chromosomal short repeats
    o_pair = rng.integers(0, npairs, n_off)
    o_qchr = rng.integers(0, nc, n_off)
    o_tchr = rng.integers(0, nc, n_off)
    o_len = rng.integers(500, 3001, n_off)
    o_qs = (rng.random(n_off) * np.maximum(L[o_qchr] - o_len - 1, 1)).astype(np.int64)
    o_ts = (rng.random(n_off) * np.maximum(L[o_tchr] - o_len - 1, 1)).astype(np.int64)
    o_ident = rng.uniform(0.80, 0.95, n_off)
    o_strand = np.where(rng.random(n_off) < 0.5, ord("+"), ord("-")).astype(np.uint8)
    pair_all = np.concatenate([pair_of, o_pair])
    qchr_all = np.concatenate([qchr, o_qchr])
    tchr_all = np.concatenate([tchr, o_tchr])
    qs_all = np.concatenate([qs, o_qs]); qe_all = np.concatenate([qe, o_qs + o_len])
    ts_all = np.concatenate([ts, o_ts]); te_all = np.concatenate([te, o_ts + o_len])
    id_all = np.concatenate([ident, o_ident])
    st_all = np.concatenate([strand, o_strand])
    # aligner-like order: by genome pair, then by query chromosome; repeats interleaved (stable)
    key = pair_all * nc + qchr_all
    order = np.argsort(key, kind="stable")
    pair_all, qchr_all, tchr_all = pair_all[order], qchr_all[order], tchr_all[order]
    qs_all, qe_all, ts_all, te_all, id_all, st_all = qs_all[order], qe_all[order], ts_all[order], te_all[order], id_all[order], st_all[order]
    ok = (qe_all <= L[qchr_all]) & (te_all <= L[tchr_all])
    qid = pa[pair_all] * nc + qchr_all
    tid = pb[pair_all] * nc + tchr_all
    ok &= qid != tid  # the filter drops self mappings anyway; keep the table free of them
    sel = np.nonzero(ok)[0]
    qid, tid = qid[sel], tid[sel]
    qs_all, qe_all, ts_all, te_all, id_all, st_all = qs_all[sel], qe_all[sel], ts_all[sel], te_all[sel], id_all[sel], st_all[sel]
    blk = np.maximum(qe_all - qs_all, te_all - ts_all)
    matches = np.rint(id_all * blk).astype(np.int64)
    identity = matches / np.maximum(blk, 1)  # what the parser derives from cols 10/11
    seq_P = np.repeat(np.arange(ng, dtype=np.uint32), nc)
    names = [f"{g}#{c}" for g in genomes for c in chroms] if with_names else None
    return Table(qid, tid, qs_all, qe_all, ts_all, te_all, blk, matches, identity, st_all, seq_P, seq_P.copy(), None, names)


def yeast_like(n_records=30000, seed=1):
    """configs[0]/[1] stand-in: 8 yeast genomes x 17 chromosomes, ~30 k records."""
    return pangenome(YEAST_GENOMES, YEAST_CHROMS, YEAST_LENGTHS, n_records, seed, np.log(8000.0), 1.0, (200, 200000),
                     frac_offdiag=0.08)


def pansn(n_records=20_000_000, seed=3, n_hap=90, with_names=False):
    """configs[2]/[3]: 90 haplotypes (S001#1..S045#2) x 24 human chromosomes."""
    genomes = [f"S{(h // 2) + 1:03d}#{(h % 2) + 1}" for h in range(n_hap)]
    return pangenome(genomes, HUMAN_CHROMS, HUMAN_LENGTHS, n_records, seed, np.log(20000.0), 1.1, (500, 1_000_000),
                     self_genome=False, with_names=with_names)


def pansn_unit_plan(n_records, n_hap=90):
    """configs[3] as ONE table made of independent genome-pair units: (ordered haplotype pairs, records drawn per unit).
    Unit u of the table is generated from the seed sequence (seed, u) alone, so a rank can build exactly its own shard
    of the 200 M-record table; the table is the concatenation of the units in ascending u."""
    pairs = [(a, b) for a in range(n_hap) for b in range(n_hap) if a != b]
    quota = np.full(len(pairs), n_records // len(pairs), np.int64)
    quota[: n_records % len(pairs)] += 1
    return pairs, quota


def pansn_units(unit_ids, n_records, seed, n_hap=90):
    """The units `unit_ids` (ascending) of the unit-structured PanSN table -> (Table, unit_sizes after the generator's own
    validity filter, in the order of unit_ids)."""
    from .table import concat
    pairs, quota = pansn_unit_plan(n_records, n_hap)
    genomes = [f"S{(h // 2) + 1:03d}#{(h % 2) + 1}" for h in range(n_hap)]
    parts, sizes = [], []
    for u in unit_ids:
        t = pangenome(genomes, HUMAN_CHROMS, HUMAN_LENGTHS, int(quota[u]), (int(seed), int(u)), np.log(20000.0), 1.1, (500, 1_000_000),
                      self_genome=False, with_names=False, pairs=[pairs[u]])
        parts.append(t)
        sizes.append(t.n)
    return concat(parts), np.asarray(sizes, np.int64)


def skew(n_pile=50_000_000, n_tiny_groups=100_000, seed=5, window=6_000_000):
    """configs[4]: one chromosome pair holding a centromeric pile + many tiny groups."""
    rng = np.random.default_rng(seed)
    n_in = int(n_pile * 0.95)
    n_out = n_pile - n_in
    chrL = 248_956_422
    w0 = 120_000_000
    ln = rng.integers(300, 5001, n_pile)
    qs = np.concatenate([w0 + (rng.random(n_in) * (window - 5000)).astype(np.int64),
                         (rng.random(n_out) * (chrL - 5001)).astype(np.int64)])
    diag = rng.random(n_pile) < 0.5
    ts = np.where(diag, qs + rng.integers(-2000, 2001, n_pile),
                  np.where(np.arange(n_pile) < n_in, w0 + (rng.random(n_pile) * (window - 5000)).astype(np.int64),
                           (rng.random(n_pile) * (chrL - 5001)).astype(np.int64)))
    ts = np.clip(ts, 0, chrL - 5001)
    ident = rng.uniform(0.70, 0.99, n_pile)
    strand = np.where(rng.random(n_pile) < 0.5, ord("+"), ord("-")).astype(np.uint8)
    # tiny groups: distinct (query,target) sequence pairs over a pool of contigs
    pool = int(np.ceil(np.sqrt(n_tiny_groups))) + 1
    gsz = rng.integers(1, 21, n_tiny_groups)
    gq = (np.arange(n_tiny_groups) // pool).astype(np.int64)
    gt = (np.arange(n_tiny_groups) % pool).astype(np.int64)
    rep = np.repeat(np.arange(n_tiny_groups), gsz)
    nt = rep.shape[0]
    t_len = rng.integers(500, 20001, nt)
    seg_start = np.concatenate(([0], np.cumsum(gsz)[:-1]))
    t_off = _segment_cumsum(t_len + rng.integers(0, 3000, nt), seg_start, rep)
    t_qs = 1000 + t_off
    t_ts = 5000 + t_off + rng.integers(-50, 51, nt)
    # sequences: 0 = CHM13#1#chr1, 1 = HG002#1#chr1, then 2*pool contigs
    qid = np.concatenate([np.zeros(n_pile, np.int64), 2 + gq[rep]])
    tid = np.concatenate([np.ones(n_pile, np.int64), 2 + pool + gt[rep]])
    qs_all = np.concatenate([qs, t_qs]); qe_all = np.concatenate([qs + ln, t_qs + t_len])
    ts_all = np.concatenate([ts, t_ts]); te_all = np.concatenate([ts + ln, t_ts + t_len])
    id_all = np.concatenate([ident, np.clip(1.0 - rng.exponential(0.01, nt), 0.7, 0.999999)])
    st_all = np.concatenate([strand, np.full(nt, ord("+"), np.uint8)])
    perm = rng.permutation(qid.shape[0]) if qid.shape[0] < 5_000_000 else np.arange(qid.shape[0])
    blk = np.maximum(qe_all - qs_all, te_all - ts_all)
    matches = np.rint(id_all * blk).astype(np.int64)
    identity = matches / np.maximum(blk, 1)
    names = ["CHM13#1#chr1", "HG002#1#chr1"] + [f"ctgA{i}" for i in range(pool)] + [f"ctgB{i}" for i in range(pool)]
    P, P2 = prefix_ids(names)
    g = lambda a: a[perm]
    return Table(g(qid), g(tid), g(qs_all), g(qe_all), g(ts_all), g(te_all), g(blk), g(matches), g(identity), g(st_all),
                        P, P2, None, names)


def write_paf_fast(table: Table, path: str):
    """The same columns as write_paf plus a cg:Z: tag, built with vectorised numpy string ops (≈ 3 s per million
    lines instead of a Python loop) — for file-level benchmarks."""
    names = np.array(table.names)
    t = table
    cols = [names[t.query_id], (t.query_end.astype(np.int64) + 1000).astype(str), t.query_start.astype(str), t.query_end.astype(str),
            np.where(t.strand == ord("+"), "+", "-"), names[t.target_id], (t.target_end.astype(np.int64) + 1000).astype(str),
            t.target_start.astype(str), t.target_end.astype(str), t.matches.astype(str), t.block_length.astype(str), np.full(t.n, "60")]
    lines = cols[0]
    for c in cols[1:]:
        lines = np.char.add(np.char.add(lines, "\t"), c)
    tags = np.char.add(np.char.add(np.char.add("\tcg:Z:", t.matches.astype(str)), "="),
                       np.char.add((t.block_length.astype(np.int64) - t.matches).astype(str), "X"))
    with open(path, "w") as f:
        f.write("\n".join(np.char.add(lines, tags).tolist()) + "\n")


def write_paf(table: Table, path: str, tags=True, seq_lengths=None):
    """Emit the table as PAF text (for the parse + filter + write path).  cols 10/11 reproduce
    matches/block_length, so the parser re-derives the same identity."""
    names = table.names
    assert names is not None, "table has no names"
    with open(path, "w") as f:
        for i in range(table.n):
            q, t = int(table.query_id[i]), int(table.target_id[i])
            m, b = int(table.matches[i]), int(table.block_length[i])
            ql = seq_lengths[q] if seq_lengths is not None else int(table.query_end[i]) + 1000
            tl = seq_lengths[t] if seq_lengths is not None else int(table.target_end[i]) + 1000
            line = (f"{names[q]}\t{ql}\t{int(table.query_start[i])}\t{int(table.query_end[i])}\t{chr(table.strand[i])}\t"
                    f"{names[t]}\t{tl}\t{int(table.target_start[i])}\t{int(table.target_end[i])}\t{m}\t{b}\t60")
            if tags:
                line += f"\tdv:f:{1.0 - m / max(b, 1):.6f}\tcg:Z:{m}={b - m}X"
            f.write(line + "\n")
