set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest5.log; tail -6 gpurun_out/r2_pytest5.log
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; tail -c 300 gpurun_out/r2_bench5.json; tail -5 gpurun_out/r2_bench5.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench5_2gpu.json 2> gpurun_out/r2_bench5_2gpu.err; tail -c 1800 gpurun_out/r2_bench5_2gpu.json; tail -3 gpurun_out/r2_bench5_2gpu.err
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes5.txt 2>&1; cat gpurun_out/r2_modes5.txt
timeout 600 python profiles/bench_modes.py --shuffle > gpurun_out/r2_modes5_shuffled.txt 2>&1; cat gpurun_out/r2_modes5_shuffled.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches5.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu5.log 2>&1
