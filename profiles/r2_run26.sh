# round 2: tree sweep for piles — tests, 1:1 on piles, modes
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest26.log; tail -8 gpurun_out/r2_pytest26.log
timeout 300 python profiles/run_pile_1to1.py 1000000 2>&1 | tail -2
timeout 300 python profiles/run_pile_1to1.py 5000000 2>&1 | tail -2
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes26.txt 2>&1; cat gpurun_out/r2_modes26.txt
