# what the driver runs at round end, plus the ncu launch list of one short bench run
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; tail -c 3000 gpurun_out/bench_r1_final2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final4.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --cpu-sample 200000 > gpurun_out/bench_under_ncu2.log 2>&1
tail -c 400 gpurun_out/bench_under_ncu2.log
