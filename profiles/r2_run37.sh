set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest37.log; tail -4 gpurun_out/r2_pytest37.log
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes37.txt 2>&1; cat gpurun_out/r2_modes37.txt
timeout 600 python profiles/bench_modes.py --shuffle --only-defaults 2>&1 | tail -1
timeout 300 python profiles/run_anchor.py
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench37.json 2> gpurun_out/r2_bench37.err; tail -2 gpurun_out/r2_bench37.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench37.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline'], d['parity'])
PY
