# round 2: group sort, groups already in order are final after the scatter — tests, bench, launch list, modes (grouped + shuffled rows), ncu of the top kernels
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest14.log; tail -8 gpurun_out/r2_pytest14.log
timeout 600 python bench.py > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; head -c 600 gpurun_out/r2_bench14.json; tail -5 gpurun_out/r2_bench14.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches14.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu14.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches14.csv > gpurun_out/r2_step_launches14.txt; python profiles/step_launches.py gpurun_out/r2_launches14.csv --agg > gpurun_out/r2_step_agg14.txt; cat gpurun_out/r2_step_agg14.txt
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes14.txt 2>&1; cat gpurun_out/r2_modes14.txt
timeout 600 python profiles/bench_modes.py --shuffle > gpurun_out/r2_modes14_shuffled.txt 2>&1; cat gpurun_out/r2_modes14_shuffled.txt
