import sys; sys.path.insert(0,'.')
import sweepga_b200 as swg
from sweepga_b200 import synth
t = synth.pansn(20_000_000, seed=3)
ctx = swg.Context(0)
dev, dres = ctx.upload(t)
cfg = swg.FilterConfig.from_cli(num_mappings="1:1", scaffold_jump="0")
for _ in range(3): st = ctx.filter_device(cfg, dev, dres)
print(st.ms_device)
