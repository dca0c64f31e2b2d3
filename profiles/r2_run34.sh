set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sweep_fuzz_gpu.py tests/test_reference_vectors.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes34.txt 2>&1; cat gpurun_out/r2_modes34.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches34_1to1.csv python profiles/run_mode_1to1.py full > gpurun_out/r2_1to1_under_ncu34.log 2>&1
python profiles/aggregate_launches.py gpurun_out/r2_launches34_1to1.csv 12 > gpurun_out/r2_agg34_1to1.txt; cat gpurun_out/r2_agg34_1to1.txt
