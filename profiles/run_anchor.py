#!/usr/bin/env python
"""The per-GPU workload of the N > 1 bench runs (shard 0 of the 8-GPU partition of configs[3], --scaffold-dist 100k) alone on one
GPU: device-resident step time; with SWG_STAGE_TIMING=1 / under ncu for the launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import sweepga_b200 as swg
wt, desc = bench.workload(8, 0, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t = swg.MappingTable(wt.query_id, wt.target_id, wt.query_start, wt.query_end, wt.target_start, wt.target_end, wt.block_length, wt.matches,
                     None, wt.strand, wt.seq_genome_id, wt.seq_genome2_id)
cfg = swg.FilterConfig.from_cli(scaffold_dist="100k")
ctx = swg.Context(0)
d_in, d_res = ctx.upload(t)
for _ in range(2):
    ctx.filter_device(cfg, d_in, d_res)
sts = [ctx.filter_device(cfg, d_in, d_res) for _ in range(3)]
print(f"anchor: {t.n} records, {min(s.ms_device for s in sts):.3f} ms, launches {sts[-1].gpu_launches}, lsd passes {sts[-1].n_sort_passes}, "
      f"groups ordered after the scatter {sts[-1].n_unsorted_groups}, kept {sts[-1].n_kept}")
