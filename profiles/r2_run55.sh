# chromosome-scale collinear groups: fixed point vs claims + sequential resolve
set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
timeout 300 python profiles/bench_large_groups.py 50000 24 2>&1 | grep "large groups\|identical\|fixpoint\] \(target\|[0-9]\)\|rror\|stages\] prefilter" | tail -8 | cut -c1-600
timeout 300 python profiles/bench_large_groups.py 200000 4 2>&1 | grep "large groups\|identical\|fixpoint\] \(target\|[0-9]\)\|rror" | tail -6 | cut -c1-600
