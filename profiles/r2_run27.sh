# round 2: bench line of the current build, launch list, ncu --set full of the kernels of the default step
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench27.json 2> gpurun_out/r2_bench27.err; tail -c 2500 gpurun_out/r2_bench27.json; tail -3 gpurun_out/r2_bench27.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench27_ref.json 2> gpurun_out/r2_bench27_ref.err; tail -c 700 gpurun_out/r2_bench27_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches27.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu27.log 2>&1
python profiles/step_launches.py gpurun_out/r2_launches27.csv > gpurun_out/r2_step_launches27.txt; python profiles/step_launches.py gpurun_out/r2_launches27.csv --agg > gpurun_out/r2_step_agg27.txt; head -24 gpurun_out/r2_step_agg27.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_prefilter|k_chain_candidates|k_chain_aggregate|k_chain_number|k_gs_|sc_flags|t_assign|t_gather' -s 20 -c 18 -o gpurun_out/r2_prof27 -f python bench.py --steps 1 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_ncu_full27.log 2>&1
ls -la gpurun_out/r2_prof27.ncu-rep
