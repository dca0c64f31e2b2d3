# round 2: sweep kernel under ncu, 1:1 on the pile, configs[4] at 5 M / 20 M / 50 M
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_flat1|sc_segmax|t_sweep_gather' -s 4 -c 6 -o gpurun_out/r2_prof25_sweep -f python profiles/run_mode_1to1.py > gpurun_out/r2_ncu25.log 2>&1
timeout 600 python profiles/run_pile_1to1.py 1000000 2>&1 | tail -2
timeout 600 python profiles/run_pile_1to1.py 5000000 2>&1 | tail -2
SWG_STAGE_TIMING=1 timeout 300 python profiles/bench_skew.py 5000000 100000 2>&1 | grep "skew\|stages\] prefilter" | tail -3 | cut -c1-900
SWG_STAGE_TIMING=1 timeout 600 python profiles/bench_skew.py 20000000 100000 2>&1 | grep "skew\|stages\] prefilter" | tail -2 | cut -c1-900
SWG_STAGE_TIMING=1 SWG_FIXPOINT_VERIFY=1 timeout 900 python profiles/bench_skew.py 50000000 100000 2>&1 | grep "skew\|stages\] prefilter\|rror" | tail -2 | cut -c1-900
