# ncu --set full of the final fixed-point kernels: k_fx_recompute<true> and k_fx_check_warp, round 3 of the 5 M pile
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fx_recompute|k_fx_check_warp" -s 6 -c 2 -f -o gpurun_out/r2_prof67_fx python profiles/bench_skew.py 5000000 100000 2>&1 | tail -2
