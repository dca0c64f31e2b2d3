# last build: the huge-group paths once more (parity), a large-jump case
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fixpoint or skew or dense or inversion or pile or edge" 2>&1 | tail -3
