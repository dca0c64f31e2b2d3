set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest28.log; tail -8 gpurun_out/r2_pytest28.log
timeout 300 python profiles/run_anchor.py
