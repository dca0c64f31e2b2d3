# final build: bench with the skew leg at 20 M (warmed, verified)
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 200 python bench.py --skew-pile 20000000 --paf-lines 0 --no-anchor --steps 5 --warmup 3 > gpurun_out/r2_bench70.json 2> gpurun_out/r2_bench70.err; tail -2 gpurun_out/r2_bench70.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench70.json').read())
print(d['value'], d['ms_per_step'], d['parity']['ok'], d.get('skew'))
PY
