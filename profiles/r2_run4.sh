set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest4.log; tail -6 gpurun_out/r2_pytest4.log
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; tail -c 300 gpurun_out/r2_bench4.json; tail -5 gpurun_out/r2_bench4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches4.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu4.log 2>&1
