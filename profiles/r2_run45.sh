# target-bucket order: the 20 M and 50 M piles with the per-round log (full), 50 M with the verification pass
set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
timeout 300 python profiles/bench_skew.py 20000000 100000 > gpurun_out/r2_fx_buckets45_20m.txt 2>&1
SWG_FIXPOINT_VERIFY=1 timeout 600 python profiles/bench_skew.py 50000000 100000 > gpurun_out/r2_fx_buckets45_50m.txt 2>&1
grep "skew\|rror\|verification\|stages\] prefilter" gpurun_out/r2_fx_buckets45_*.txt | cut -c1-700
