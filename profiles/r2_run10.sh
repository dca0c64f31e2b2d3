# round 2: ncu --set full of the shared-memory-mask ranking variant of the sort pass (sort_bench, random keys then grouped keys)
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 7 -c 7 -o gpurun_out/r2_prof10_random -f build/sort_bench_r2 20000000 1 > gpurun_out/r2_ncu10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 21 -c 7 -o gpurun_out/r2_prof10_grouped -f build/sort_bench_r2 20000000 1 >> gpurun_out/r2_ncu10.log 2>&1
ls -la gpurun_out/*.ncu-rep
