# round 2, second GPU call: tests, bench, stage timing of the 20 M step, launch list, ncu --set full of the main kernels
set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest2.log; tail -8 gpurun_out/r2_pytest2.log
timeout 600 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 600 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
SWG_STAGE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > /dev/null 2> gpurun_out/r2_stages2.txt; grep -m3 "swg stages" gpurun_out/r2_stages2.txt | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches2.csv python bench.py --steps 2 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_bench_under_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rs_onesweep_kernel|k_chain_number|k_chain_candidates|k_prefilter|k_chain_aggregate|k_chain_resolve_warp' -s 22 -c 18 -o gpurun_out/r2_prof2 python bench.py --steps 1 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_ncu_full2.log 2>&1
ls -la gpurun_out/r2_prof2.ncu-rep
