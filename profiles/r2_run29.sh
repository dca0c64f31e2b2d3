set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_properties_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench29.json 2> gpurun_out/r2_bench29.err; tail -3 gpurun_out/r2_bench29.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench29.json').read())
print(d['ms_per_step'], d['e2e'])
PY
