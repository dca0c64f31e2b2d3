# full GPU suite, smoke, default bench (skew leg at 50 M)
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r2_bench66.json 2> gpurun_out/r2_bench66.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2_bench66.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench66.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['parity'], d.get('skew'))
PY
