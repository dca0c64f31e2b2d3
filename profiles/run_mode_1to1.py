"""One mode of the 20 M table alone (for ncu launch lists): python profiles/run_mode_1to1.py [full]   (default: 1:1, no scaffolding;
full: 1:1 / 1:1 with scaffolding)."""
import sys; sys.path.insert(0,'.')
import sweepga_b200 as swg
from sweepga_b200 import synth
t = synth.pansn(20_000_000, seed=3)
ctx = swg.Context(0)
dev, dres = ctx.upload(t)
cfg = swg.FilterConfig.from_cli(num_mappings="1:1", scaffold_filter="1:1") if "full" in sys.argv else swg.FilterConfig.from_cli(num_mappings="1:1", scaffold_jump="0")
for _ in range(2): st = ctx.filter_device(cfg, dev, dres)
print(st.ms_device, st.gpu_launches)
