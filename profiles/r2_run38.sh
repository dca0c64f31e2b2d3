# round 2: k_prefilter with 1 / 2 / 4 records per thread (build/variants/pf*.so)
cd /root/repo; mkdir -p gpurun_out
cp sweepga_b200/libsweepga_b200.so /tmp/lib_keep.so
for v in pf1 pf2 pf4; do
  cp build/variants/$v.so sweepga_b200/libsweepga_b200.so
  echo "== $v"
  timeout 300 python bench.py --steps 10 --warmup 3 --paf-lines 0 --skew-pile 0 --no-anchor --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms_per_step', round(d['ms_per_step'],3), 'prefilter_ms', round(r['launch_ms'],4), 'frac', round(r['frac'],3), 'bytes', r['bytes_per_record'])"
done 2>&1 | tee gpurun_out/r2_prefilter_items.txt
cp /tmp/lib_keep.so sweepga_b200/libsweepga_b200.so
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_properties_gpu.py -m gpu -x -q -k "yeast_configs or pansn or edge or range or fuzz_dense or determinism or idempot" 2>&1 | tail -3
