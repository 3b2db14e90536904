import csv,sys,collections,subprocess
rep=sys.argv[1]; tiles=int(sys.argv[2]) if len(sys.argv)>2 else 50040
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; d=dict(zip(hdr,rows[2]))
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__bytes_read.sum.per_second','launch__grid_size','launch__registers_per_thread','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys: print(k, d.get(k))
for k,v in d.items():
    if 'issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k: print(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v)
print('instr/tile', float(d['smsp__inst_executed.sum'])/tiles, 'per pt-group', float(d['smsp__inst_executed.sum'])/tiles/128)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
byop=collections.Counter(); wf=collections.Counter(); samp=collections.Counter()
for r in rows[2:]:
    if len(r)<len(hdr): continue
    s=r[ix['Source']].strip()
    t=s.split()
    op=t[1] if t[0].startswith('@') else t[0]
    op='.'.join(op.split('.')[:2]) if op.startswith(('ATOMS','LDG','LDS','STS','UBLKCP','SYNCS','BAR')) else op.split('.')[0]
    n=int(r[ix['Instructions Executed']])
    byop[op]+=n; wf[op]+=int(r[ix['L1 Wavefronts Shared']]); samp[op]+=int(r[ix['# Samples']])
tot=sum(samp.values())
for op,n in byop.most_common(28):
    print(f'{op:18s} {n/tiles:8.1f}/tile {n/tiles/128:6.2f}/pt smem_wf/tile {wf[op]/tiles:7.1f} samples {100*samp[op]/tot:5.1f}%')
