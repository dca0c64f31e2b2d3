# A/B of radix-sort digit widths / tile shapes (build/variants/s*.so, made by hand with -DSWG_RS_BITS=.. -DSWG_RS_THREADS=.. -DSWG_RS_ITEMS=..)
for v in ${1:-s0 s1 s2}; do
  cp build/variants/$v.so sweepga_b200/libsweepga_b200.so
  echo "== $v"
  if [ "$v" != "s0" ]; then timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "yeast_configs or pansn_400k or wide_key" 2>&1 | tail -1; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --paf-lines 0 --skew-pile 0 --cpu-sample 100000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms_per_step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), 'launches/step', d['gpu_launches']/10, 'pass_ms', round(r['launch_ms'],4), 'frac', round(r['frac'],3), 'kept', d['config']['stats']['n_kept'])"
done
