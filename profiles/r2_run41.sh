# round 2: last check of the committed build — all GPU tests, smoke(), default bench
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest41.log; tail -3 gpurun_out/r2_pytest41.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench41.json 2> gpurun_out/r2_bench41.err; tail -2 gpurun_out/r2_bench41.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench41.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity']['ok'], d['scale_anchor']['ms_per_step'], d['modes'])
PY
