# query-bucket width of the inversion capture chosen from counted entries / cells: parity tests, 20 M and 50 M piles
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sweep_fuzz_gpu.py -m gpu -x -q -k "fixpoint or skew or dense or inversion or fuzz" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
timeout 600 python profiles/bench_skew.py 20000000 100000 2>&1 | grep "skew\|rror\|inversion\]\|stages\] prefilter" | cut -c1-700
timeout 600 python profiles/bench_skew.py 50000000 100000 2>&1 | grep "skew\|rror\|inversion\]\|stages\] prefilter" | cut -c1-700
