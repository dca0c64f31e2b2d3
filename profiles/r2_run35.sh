set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "group_sort or pansn or yeast_configs or tree_sweep_pile or record_sort" 2>&1 | tail -3
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes35.txt 2>&1; cat gpurun_out/r2_modes35.txt
timeout 300 python profiles/run_anchor.py
