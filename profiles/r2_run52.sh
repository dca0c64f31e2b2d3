# pruning with the bucket's nearest edge (q_gap^2 + r_min^2 > best d); bucket widths at 50 M
set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
for nw in 3 5 7; do
  echo "== SWG_FX_BUCKET_NARROW=$nw"
  SWG_FX_BUCKET_NARROW=$nw timeout 300 python profiles/bench_skew.py 50000000 100000 2>&1 | grep "round 0\|round 1:\|round 3:\|round 8:\|round 16:\|buckets\|skew\|rror\|stages\] prefilter" | cut -c1-330
done > gpurun_out/r2_fx_buckets52.txt 2>&1
cat gpurun_out/r2_fx_buckets52.txt | grep -v "^+" | cut -c1-260
