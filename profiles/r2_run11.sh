# round 2: group sort (counting sort by group + per-group ordering) — tests, bench, stages, launch list, shuffled modes
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest11.log; tail -12 gpurun_out/r2_pytest11.log
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; head -c 1800 gpurun_out/r2_bench11.json; tail -5 gpurun_out/r2_bench11.err
SWG_STAGE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > /dev/null 2> gpurun_out/r2_stages11.txt; grep "swg stages" gpurun_out/r2_stages11.txt | sed -n 3p
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches11.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu11.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches11.csv > gpurun_out/r2_step_launches11.txt; python profiles/step_launches.py gpurun_out/r2_launches11.csv --agg > gpurun_out/r2_step_agg11.txt; cat gpurun_out/r2_step_agg11.txt
timeout 600 python profiles/bench_modes.py --shuffle > gpurun_out/r2_modes11_shuffled.txt 2>&1; cat gpurun_out/r2_modes11_shuffled.txt
SWG_NO_GROUP_SORT=1 timeout 600 python profiles/bench_modes.py --shuffle > gpurun_out/r2_modes11_shuffled_lsd.txt 2>&1; cat gpurun_out/r2_modes11_shuffled_lsd.txt
