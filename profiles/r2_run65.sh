set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
timeout 300 python profiles/bench_large_groups.py 50000 24 2>&1 | grep "large groups\|identical\|rror\|fixpoint\]\|stages\] prefilter\|swg inversion" | head -9 | cut -c1-500
