# round 2: ncu --set full of the main kernels of the 20 M default step (second step; -lineinfo build)
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rs_onesweep_kernel|rs_histogram|k_chain_number|k_chain_candidates|k_prefilter|k_chain_aggregate|sc_flags|t_assign|t_gather' -s 25 -c 24 -o gpurun_out/r2_prof8 -f python bench.py --steps 1 --warmup 1 --paf-lines 0 --skew-pile 0 --no-parity --no-anchor > gpurun_out/r2_ncu_full8.log 2>&1
ls -la gpurun_out/r2_prof8.ncu-rep
ncu -i gpurun_out/r2_prof8.ncu-rep --page raw --csv > gpurun_out/r2_prof8_raw.csv 2>/dev/null; wc -c gpurun_out/r2_prof8_raw.csv
