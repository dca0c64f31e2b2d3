# target-bucket order of the fixed-point searches: parity tests, then configs[4] piles in both modes
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense" 2>&1 | tail -5
export SWG_STAGE_TIMING=1
for n in 5000000 20000000; do
  timeout 300 python profiles/bench_skew.py $n 100000 2>&1 | grep "fixpoint\]\|skew\|rror\|stages\] prefilter" | tail -40 | cut -c1-600
done > gpurun_out/r2_fx_buckets44.txt 2>&1
SWG_FX_NO_BUCKETS=1 timeout 300 python profiles/bench_skew.py 5000000 100000 2>&1 | grep "skew\|rror" | tail -3 >> gpurun_out/r2_fx_buckets44.txt
tail -60 gpurun_out/r2_fx_buckets44.txt
