set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "group_sort or pansn or yeast_configs or record_sort or skew" 2>&1 | tail -6
timeout 300 python profiles/run_anchor.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches19_anchor.csv python profiles/run_anchor.py > gpurun_out/r2_anchor_under_ncu19.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches19_anchor.csv --agg > gpurun_out/r2_step_agg19_anchor.txt; head -12 gpurun_out/r2_step_agg19_anchor.txt
timeout 600 python profiles/bench_modes.py --only-defaults
