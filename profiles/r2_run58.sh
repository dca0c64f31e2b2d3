set -x
cd /root/repo; mkdir -p gpurun_out
export SWG_STAGE_TIMING=1
timeout 600 python profiles/bench_skew.py 20000000 100000 2>&1 | grep "skew\|rror\|inversion\]" | cut -c1-700
