# round 2, first GPU call: tests, bench with the parity gate, CUB yardstick, sort variants, stage timing, launch list
set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest1.log; tail -8 gpurun_out/r2_pytest1.log
timeout 600 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 1500 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/cub_yardstick.cu -o /tmp/cuby 2>&1 | tail -2; timeout 120 /tmp/cuby > gpurun_out/r2_cub_yardstick.txt 2>&1; cat gpurun_out/r2_cub_yardstick.txt
timeout 900 python profiles/tune_sort.py 384x14 384x14x8x2x3 384x16x8x2x2 384x18x8x2x2 512x12x8x2x2 384x12x8x2x3 256x16x8x2x4 > gpurun_out/r2_tune_sort1.txt 2>&1; cat gpurun_out/r2_tune_sort1.txt
SWG_STAGE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor 2>&1 | grep "swg stages" | tail -2 > gpurun_out/r2_stages1.txt; cat gpurun_out/r2_stages1.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches1.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu1.log 2>&1
tail -c 300 gpurun_out/r2_bench_under_ncu1.log
