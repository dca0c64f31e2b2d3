# fixed-point chaining: parity tests, then the configs[4] piles with per-round / per-stage timing on stderr
set -x
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense" 2>&1 | tail -5
export SWG_STAGE_TIMING=1
for n in 1000000 5000000 20000000 50000000; do
  timeout 500 python profiles/bench_skew.py $n 100000 2>&1 | grep "fixpoint\]\|skew\|rror\|stages\] prefilter" | tail -90 | cut -c1-1500
done
