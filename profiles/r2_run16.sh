# round 2: chunked run ranges, early LSD choice for ungrouped rows — tests, bench (with anchor), launch lists, modes
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest16.log; tail -8 gpurun_out/r2_pytest16.log
SWG_STAGE_TIMING=1 timeout 600 python bench.py --paf-lines 0 --skew-pile 0 > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err; grep "group sort" gpurun_out/r2_bench16.err | sort | uniq -c
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench16.json').read())
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['bytes_per_record'], d['detail']['stats'], d['scale_anchor'], d['parity'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches16.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu16.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches16.csv > gpurun_out/r2_step_launches16.txt; python profiles/step_launches.py gpurun_out/r2_launches16.csv --agg > gpurun_out/r2_step_agg16.txt; head -22 gpurun_out/r2_step_agg16.txt
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes16.txt 2>&1; cat gpurun_out/r2_modes16.txt
timeout 600 python profiles/bench_modes.py --shuffle > gpurun_out/r2_modes16_shuffled.txt 2>&1; cat gpurun_out/r2_modes16_shuffled.txt
