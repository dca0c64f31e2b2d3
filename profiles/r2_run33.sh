# round 2: the primary sweeps' sort through the group sort — tests, modes
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest33.log; tail -8 gpurun_out/r2_pytest33.log
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes33.txt 2>&1; cat gpurun_out/r2_modes33.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches33_1to1.csv python profiles/run_mode_1to1.py full > gpurun_out/r2_1to1_under_ncu33.log 2>&1
python profiles/aggregate_launches.py gpurun_out/r2_launches33_1to1.csv 30 > gpurun_out/r2_agg33_1to1.txt; cat gpurun_out/r2_agg33_1to1.txt
