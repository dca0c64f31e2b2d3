"""Seeded synthetic mapping tables of the BASELINE.json shapes as `MappingTable`s.

The generators themselves are pure numpy and live in the top-level `workloads` package (so that the CPU
reference arm of bench.py can build the same tables without loading this package's native library); this
module only re-types their results.
"""
from workloads import synth as _s
from workloads.synth import (HUMAN_CHROMS, HUMAN_LENGTHS, YEAST_CHROMS, YEAST_GENOMES, YEAST_LENGTHS, write_paf,  # noqa: F401
                             write_paf_fast)

from .api import MappingTable


def _typed(t):
    return MappingTable(t.query_id, t.target_id, t.query_start, t.query_end, t.target_start, t.target_end, t.block_length,
                        t.matches, t.identity, t.strand, t.seq_genome_id, t.seq_genome2_id, t.score, t.names, t.rank)


def pangenome(*a, **k):
    return _typed(_s.pangenome(*a, **k))


def yeast_like(*a, **k):
    return _typed(_s.yeast_like(*a, **k))


def pansn(*a, **k):
    return _typed(_s.pansn(*a, **k))


def skew(*a, **k):
    return _typed(_s.skew(*a, **k))
