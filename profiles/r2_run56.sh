# finer diagonal buckets of the inversion capture; force-collect tests; 50 M pile
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense or inversion" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
timeout 600 python profiles/bench_skew.py 50000000 100000 2>&1 | grep "skew\|rror\|stages\] prefilter" | cut -c1-700
timeout 600 python profiles/bench_skew.py 5000000 100000 2>&1 | grep "skew\|rror\|stages\] prefilter" | tail -2 | cut -c1-700
