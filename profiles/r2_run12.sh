# round 2: group sort with run-based slots — tests, bench, launch list, shuffled modes
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest12.log; tail -12 gpurun_out/r2_pytest12.log
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; head -c 600 gpurun_out/r2_bench12.json; tail -5 gpurun_out/r2_bench12.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches12.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu12.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches12.csv > gpurun_out/r2_step_launches12.txt; python profiles/step_launches.py gpurun_out/r2_launches12.csv --agg > gpurun_out/r2_step_agg12.txt; cat gpurun_out/r2_step_agg12.txt
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes12.txt 2>&1; cat gpurun_out/r2_modes12.txt
timeout 600 python profiles/bench_modes.py --shuffle > gpurun_out/r2_modes12_shuffled.txt 2>&1; cat gpurun_out/r2_modes12_shuffled.txt
