set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest6.log; tail -6 gpurun_out/r2_pytest6.log
timeout 600 python bench.py --paf-lines 0 --no-anchor > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; tail -c 900 gpurun_out/r2_bench6.json; tail -5 gpurun_out/r2_bench6.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches6.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu6.log 2>&1
