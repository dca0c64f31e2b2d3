# round 2: group sort, groups already in order are final after the scatter — tests, bench, launch list, modes (grouped + shuffled rows), ncu of the top kernels
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest13.log; tail -8 gpurun_out/r2_pytest13.log
timeout 600 python bench.py > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; head -c 600 gpurun_out/r2_bench13.json; tail -5 gpurun_out/r2_bench13.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches13.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu13.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches13.csv > gpurun_out/r2_step_launches13.txt; python profiles/step_launches.py gpurun_out/r2_launches13.csv --agg > gpurun_out/r2_step_agg13.txt; cat gpurun_out/r2_step_agg13.txt
timeout 600 python profiles/bench_modes.py > gpurun_out/r2_modes13.txt 2>&1; cat gpurun_out/r2_modes13.txt
timeout 600 python profiles/bench_modes.py --shuffle > gpurun_out/r2_modes13_shuffled.txt 2>&1; cat gpurun_out/r2_modes13_shuffled.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_prefilter|k_chain_candidates|k_chain_aggregate|k_chain_number|k_gs_|sc_flags|t_assign|t_gather' -s 20 -c 16 -o gpurun_out/r2_prof13 -f python bench.py --steps 1 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_ncu_full13.log 2>&1
ls -la gpurun_out/r2_prof13.ncu-rep
