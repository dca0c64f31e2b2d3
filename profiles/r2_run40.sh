# round 2: target-bucket pre-test in the outward rounds of the fixed-point search — parity, then the piles
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense or inversion_grid_yeast" 2>&1 | tail -3
SWG_STAGE_TIMING=1 timeout 300 python profiles/bench_skew.py 5000000 100000 2>&1 | grep "skew\|swg fixpoint\]" | tail -3 | cut -c1-300
SWG_STAGE_TIMING=1 timeout 600 python profiles/bench_skew.py 20000000 100000 2>&1 | grep "skew\|swg fixpoint\]\|stages\] prefilter" | tail -3 | cut -c1-700
