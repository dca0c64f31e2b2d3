# round 2, third GPU call (2 GPUs): all GPU tests incl. the multi-GPU ones, 1-GPU bench, 2-GPU bench (both arms), launch list
set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest3.log; tail -6 gpurun_out/r2_pytest3.log
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -c 300 gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench3_2gpu.json 2> gpurun_out/r2_bench3_2gpu.err; tail -c 2500 gpurun_out/r2_bench3_2gpu.json; tail -5 gpurun_out/r2_bench3_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench3_2gpu_ref.json 2> gpurun_out/r2_bench3_2gpu_ref.err; tail -c 800 gpurun_out/r2_bench3_2gpu_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches3.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu3.log 2>&1
