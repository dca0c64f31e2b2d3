# snapshot summaries off the sorted lists (no atomicMin passes); bucket widths at 50 M again
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
for nw in 5 4 6; do
  echo "== SWG_FX_BUCKET_NARROW=$nw"
  SWG_FX_BUCKET_NARROW=$nw timeout 300 python profiles/bench_skew.py 50000000 100000 2>&1 | grep "round 0\|round 1:\|round 3:\|round 8:\|round 16:\|round 40:\|buckets\|skew\|rror\|stages\] prefilter" | cut -c1-330
done > gpurun_out/r2_fx_buckets60.txt 2>&1
grep -v "^+" gpurun_out/r2_fx_buckets60.txt | cut -c1-250
