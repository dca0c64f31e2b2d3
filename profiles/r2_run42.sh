# round 2: ncu --set full of the sweep kernels of the final build (1:1, no scaffolding, 20 M)
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sweep_flat1|sc_segmax|t_sweep_gather|k_sweep_keys' -s 8 -c 10 -o gpurun_out/r2_prof42_sweep -f python profiles/run_mode_1to1.py > gpurun_out/r2_ncu42.log 2>&1
ls -la gpurun_out/r2_prof42_sweep.ncu-rep
