# round 0 of the fixed point: thread-per-position linear scan first; parity tests, chromosome-scale groups, 50 M pile
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense or pile" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
timeout 300 python profiles/bench_large_groups.py 50000 24 2>&1 | grep "large groups\|identical\|rror\|round 0" | head -6 | cut -c1-300
timeout 300 python profiles/bench_skew.py 50000000 100000 2>&1 | grep "skew\|rror\|round 0\|stages\] prefilter" | tail -3 | cut -c1-500
