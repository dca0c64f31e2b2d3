# ncu of k_fx_recompute<true>, round 3 of the 5 M pile (bucket order)
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fx_recompute -s 3 -c 1 -f -o gpurun_out/r2_prof49_fx python profiles/bench_skew.py 5000000 100000 2>&1 | tail -3
