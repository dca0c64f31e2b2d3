# round 2, final build: all GPU tests, smoke(), bench (both arms), launch list
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest36.log; tail -4 gpurun_out/r2_pytest36.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r2_bench36.json 2> gpurun_out/r2_bench36.err; tail -3 gpurun_out/r2_bench36.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench36_ref.json 2> gpurun_out/r2_bench36_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench36.json').read())
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['parity'], d['modes'], d['small'], d['scale_anchor'], d['gpu_launches'])
r=json.loads(open('gpurun_out/r2_bench36_ref.json').read())
print(r['value'], r['oracle_1t'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches36.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu36.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches36.csv > gpurun_out/r2_step_launches36.txt; python profiles/step_launches.py gpurun_out/r2_launches36.csv --agg > gpurun_out/r2_step_agg36.txt; head -12 gpurun_out/r2_step_agg36.txt
