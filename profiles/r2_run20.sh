# round 2 (2 GPUs): multi-GPU tests, 2-GPU bench both arms
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench20_2gpu.json 2> gpurun_out/r2_bench20_2gpu.err; tail -c 2600 gpurun_out/r2_bench20_2gpu.json; tail -3 gpurun_out/r2_bench20_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench20_2gpu_ref.json 2> gpurun_out/r2_bench20_2gpu_ref.err; tail -c 600 gpurun_out/r2_bench20_2gpu_ref.json
