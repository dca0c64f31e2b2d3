# fixed-point chaining + bucketed inversion capture: parity tests, then the configs[4] piles with per-round / per-stage
# timing on stderr.  Usage: bash profiles/run_fixpoint_checks.sh "1000000 5000000 20000000"
set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense or inversion or yeast" 2>&1 | tail -5
export SWG_STAGE_TIMING=1
for n in ${1:-1000000 5000000}; do
  timeout 500 python profiles/bench_skew.py $n 100000 2>&1 | grep "fixpoint\]\|skew\|rror\|stages\] prefilter" | tail -90 | cut -c1-1500
done
