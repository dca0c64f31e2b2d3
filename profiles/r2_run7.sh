# round 2, re-entry baseline of HEAD: tests, bench, stage timing, launch list, ncu --set full of the main kernels
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest7.log; tail -6 gpurun_out/r2_pytest7.log
timeout 600 python bench.py > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; tail -c 3000 gpurun_out/r2_bench7.json; tail -5 gpurun_out/r2_bench7.err
SWG_STAGE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > /dev/null 2> gpurun_out/r2_stages7.txt; grep "swg stages" gpurun_out/r2_stages7.txt | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches7.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu7.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches7.csv > gpurun_out/r2_step_launches7.txt; python profiles/step_launches.py gpurun_out/r2_launches7.csv --agg > gpurun_out/r2_step_agg7.txt; cat gpurun_out/r2_step_agg7.txt
nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/cub_yardstick.cu -o /tmp/cuby 2>&1 | tail -2; timeout 120 /tmp/cuby > gpurun_out/r2_cub_yardstick.txt 2>&1; cat gpurun_out/r2_cub_yardstick.txt
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes7.txt 2>&1; cat gpurun_out/r2_modes7.txt
