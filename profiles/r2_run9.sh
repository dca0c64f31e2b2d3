# round 2: ranking / tile-shape variants of the record sort pass (profiles/sort_bench.cu, prebuilt under build/)
cd /root/repo; mkdir -p gpurun_out
for b in build/sort_bench_*; do echo "== $b"; timeout 120 $b 20000000 5; done 2>&1 | tee gpurun_out/r2_sort_variants9.txt
