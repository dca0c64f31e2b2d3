# round 2: where the time goes with shuffled rows; run / group counts; the scaling anchor
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches15_shuffled.csv python profiles/bench_modes.py --shuffle --only-defaults > gpurun_out/r2_shuffled_under_ncu15.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches15_shuffled.csv --agg > gpurun_out/r2_step_agg15_shuffled.txt; cat gpurun_out/r2_step_agg15_shuffled.txt
SWG_STAGE_TIMING=1 timeout 300 python profiles/bench_modes.py --only-defaults 2>&1 | grep "group sort" | tail -1
SWG_STAGE_TIMING=1 timeout 300 python profiles/bench_modes.py --shuffle --only-defaults 2>&1 | grep "group sort" | tail -1
SWG_STAGE_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; grep "group sort" gpurun_out/r2_bench15.err | tail -2; grep "swg stages" gpurun_out/r2_bench15.err | tail -1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench15.json').read())
print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['bytes_per_record'], d['scale_anchor'])
PY
