import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__cycles_elapsed.max','smsp__thread_inst_executed_per_inst_executed.ratio']
r=rows[2]
print('---',r[hdr.index('Kernel Name')][:70])
for k in keys:
    if k in hdr: print(f'  {k:75s} {r[hdr.index(k)]:>16s} {rows[1][hdr.index(k)]}')
for i,h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(r[i])>0.05: print(f'  {h:75s} {r[i]:>16s}')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]
isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); ismp=hdr.index('# Samples')
body=[]
for r in rows[2:]:
    if len(r)<10 or r[0]=='Kernel Name': break
    body.append(r)
tot=sum(int(r[ismp]) for r in body); tex=sum(int(r[iex]) for r in body)
print('total samples',tot,'inst',tex)
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
for k,r in enumerate(body):
    if int(r[ismp])>tot*thr: print(f"{k:4d} ex={int(r[iex])/1e6:7.2f}M smp={r[ismp]:>6s}  {r[isrc].strip()[:90]}")
