# 32 B snapshot (first picker's d, list range): parity tests, 50 M pile, ncu of round 3 on the 5 M pile
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
SWG_FIXPOINT_VERIFY=1 timeout 600 python profiles/bench_skew.py 50000000 100000 > gpurun_out/r2_fx_buckets51_50m.txt 2>&1
grep "skew\|rror\|verification\|stages\] prefilter" gpurun_out/r2_fx_buckets51_*.txt | cut -c1-700
unset SWG_STAGE_TIMING
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fx_recompute -s 3 -c 1 -f -o gpurun_out/r2_prof51_fx python profiles/bench_skew.py 5000000 100000 2>&1 | tail -2
