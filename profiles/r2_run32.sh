# round 2 (8 GPUs): the scaling bench as the driver runs it, both arms
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench32_8gpu.json 2> gpurun_out/r2_bench32_8gpu.err; tail -3 gpurun_out/r2_bench32_8gpu.err; tail -c 2800 gpurun_out/r2_bench32_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench32_8gpu_ref.json 2> gpurun_out/r2_bench32_8gpu_ref.err; tail -c 300 gpurun_out/r2_bench32_8gpu_ref.json
