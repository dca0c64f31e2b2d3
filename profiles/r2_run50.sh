# picker lists in position order with prefix minima (binary search instead of the walk): parity tests, 20 M and 50 M piles
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense or inversion" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
timeout 300 python profiles/bench_skew.py 20000000 100000 > gpurun_out/r2_fx_buckets50_20m.txt 2>&1
SWG_FIXPOINT_VERIFY=1 timeout 600 python profiles/bench_skew.py 50000000 100000 > gpurun_out/r2_fx_buckets50_50m.txt 2>&1
grep "skew\|rror\|verification\|stages\] prefilter" gpurun_out/r2_fx_buckets50_*.txt | cut -c1-700
