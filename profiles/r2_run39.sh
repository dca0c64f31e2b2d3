set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest39.log; tail -4 gpurun_out/r2_pytest39.log
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes39.txt 2>&1; cat gpurun_out/r2_modes39.txt
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['small'])"
