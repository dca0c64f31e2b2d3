# A/B of k_fx_recompute builds (build/variants/*.so, made by hand) on one pile: prints the stage line of the timed call
export SWG_STAGE_TIMING=1
for v in ${2:-v0 v1 v2 v3}; do
  cp build/variants/$v.so sweepga_b200/libsweepga_b200.so
  echo "== $v"
  timeout 300 python profiles/bench_skew.py ${1:-8000000} 100000 2>&1 | grep "stages\] prefilter\|^skew\|rror\|rounds," | tail -3 | cut -c1-600
done
