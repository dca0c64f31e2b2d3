# round 2 (2 GPUs): all GPU tests, 1-GPU bench, 2-GPU bench both arms
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest31.log; tail -5 gpurun_out/r2_pytest31.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench31_2gpu.json 2> gpurun_out/r2_bench31_2gpu.err; tail -3 gpurun_out/r2_bench31_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench31_2gpu.json').read())
print(d['value'], d['ms_per_step'], d['detail']['filter_only_ms_per_step'], d['e2e'], d['parity'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench31_2gpu_ref.json 2> gpurun_out/r2_bench31_2gpu_ref.err; tail -c 300 gpurun_out/r2_bench31_2gpu_ref.json
