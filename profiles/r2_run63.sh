# narrow query buckets of the inversion capture only where chains are many: parity tests, chromosome-scale collinear groups
set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "inversion or skew or pile" 2>&1 | tail -3
export SWG_STAGE_TIMING=1
timeout 300 python profiles/bench_large_groups.py 50000 24 2>&1 | grep "large groups\|identical\|rror\|stages\] prefilter\|swg inversion\|fixpoint\] \(target\|round\)" | head -14 | cut -c1-420
