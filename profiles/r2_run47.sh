# bucket width of the target-bucket order (compact 8 B entries, buckets visited outwards): 20 M pile per width, parity tests
set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
for nw in 0 2 3 4 5; do
  echo "== SWG_FX_BUCKET_NARROW=$nw"
  SWG_FX_BUCKET_NARROW=$nw timeout 300 python profiles/bench_skew.py 20000000 100000 2>&1 | grep "round 0\|round 1:\|round 3:\|round 8:\|buckets\|skew\|rror\|stages\] prefilter" | cut -c1-330
done > gpurun_out/r2_fx_buckets47.txt 2>&1
unset SWG_STAGE_TIMING
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense" 2>&1 | tail -3
