"""Synthetic workloads of the BASELINE.json shapes, as plain numpy column tables.

Pure Python + numpy on purpose: bench.py's `--impl reference` arm and the oracle binding import this
package without mapping the product library (libsweepga_b200.so) into the process.
"""
from .table import Table, concat, prefix_P, prefix_P2, prefix_ids, lpt_shards  # noqa: F401
