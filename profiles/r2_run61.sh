# warp-cooperative k_fx_check: parity tests, 20 M and 50 M piles
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
for n in 20000000 50000000; do
  timeout 300 python profiles/bench_skew.py $n 100000 2>&1 | grep "round 1:\|round 8:\|round 16:\|round 24:\|round 40:\|skew\|rror\|stages\] prefilter" | cut -c1-330
done > gpurun_out/r2_fx_buckets61.txt 2>&1
grep -v "^+" gpurun_out/r2_fx_buckets61.txt | cut -c1-250
