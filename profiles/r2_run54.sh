# chromosome-scale groups on collinear data; memcheck over the bucket-order fixed point
set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
timeout 300 python profiles/bench_large_groups.py 2000000 2 2>&1 | grep "large groups\|identical\|fixpoint\] \(target\|[0-9]\)\|rror" | tail -12
timeout 300 python profiles/bench_large_groups.py 4000000 4 2>&1 | grep "large groups\|identical\|fixpoint\] \(target\|[0-9]\)\|rror" | tail -12
unset SWG_STAGE_TIMING
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint_search_orders or inversion_grid_without" > gpurun_out/r2_sanitizer_memcheck_buckets.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_buckets.txt
