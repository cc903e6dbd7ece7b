"""One block of headline metrics per distinct kernel of an ncu report (first captured launch of each):
python profiles/ncu_summary_all.py report.ncu-rep > summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__thread_inst_executed_per_inst_executed.ratio']
seen = set()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[hdr.index('Kernel Name')]
    if name in seen:
        continue
    seen.add(name)
    print('---', name[:110])
    for k in keys:
        if k in hdr:
            print(f'  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}')
    st = [(float(r[i]), h) for i, h in enumerate(hdr)
          if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and r[i] not in ('', 'n/a')]
    for v, h in sorted(st, reverse=True)[:4]:
        print(f'  stall {h.split("issue_stalled_")[1].split("_per_issue")[0]:30s} {v:8.2f}')
