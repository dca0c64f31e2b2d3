import csv,sys,subprocess,io
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=rows[0]; units=rows[1]; data=rows[2:]
col={h:i for i,h in enumerate(hdr)}
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']
stalls=[h for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for r in data:
    print(r[col['Kernel Name']][:46])
    print('   ', ' '.join(f"{k.split('.')[0].replace('smsp__','').replace('sm__','').replace('l1tex__','')}={r[col[k]]}" for k in keys if k in col))
    st=sorted(((float(r[col[s]].replace(',','')),s) for s in stalls if r[col[s]] not in ('','n/a')),reverse=True)[:7]
    print('    stalls:', ' | '.join(f"{s.split('issue_stalled_')[1].split('_per')[0]}={v:.2f}" for v,s in st))
