set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
timeout 600 python profiles/bench_skew.py 50000000 100000 > gpurun_out/r2_fx_buckets46_50m.txt 2>&1
grep "skew\|rror\|verification\|stages\] prefilter" gpurun_out/r2_fx_buckets46_*.txt | cut -c1-700
