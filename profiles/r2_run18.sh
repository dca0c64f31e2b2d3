set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches18_anchor.csv python profiles/run_anchor.py > gpurun_out/r2_anchor_under_ncu18.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches18_anchor.csv --agg > gpurun_out/r2_step_agg18_anchor.txt; head -30 gpurun_out/r2_step_agg18_anchor.txt
