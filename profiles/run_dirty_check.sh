# the dirty-block skip of k_fx_check (default on; SWG_FX_NO_DIRTY=1 checks every position every round): parity + timing
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew" 2>&1 | tail -2
export SWG_STAGE_TIMING=1 SWG_FIXPOINT_VERIFY=1
for n in ${1:-20000000}; do
  timeout 300 python profiles/bench_skew.py $n 100000 2>&1 | grep "verification\|stages\] prefilter\|^skew\|rror\|rounds," | tail -4 | cut -c1-500
done
export SWG_FX_NO_DIRTY=1
timeout 300 python profiles/bench_skew.py ${1:-20000000} 100000 2>&1 | grep "stages\] prefilter\|^skew" | tail -2 | cut -c1-500
