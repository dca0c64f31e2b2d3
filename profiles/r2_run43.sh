set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench43.json 2> gpurun_out/r2_bench43.err; tail -2 gpurun_out/r2_bench43.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench43.json').read())
print(d['value'], d['modes'].keys(), d['modes']['1:1/1:1 on the 20 M table'])
PY
