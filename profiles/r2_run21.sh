set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_properties_gpu.py tests/test_scores_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "yeast_configs or edge or range" 2>&1 | tail -3
timeout 600 python bench.py --paf-lines 0 --skew-pile 0 --no-anchor > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err; tail -3 gpurun_out/r2_bench21.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench21.json').read())
print(d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['parity'])
PY
