"""Dump every SASS instruction of the profiled kernel with its executed count and samples: python ncu_sass.py rep"""
import csv, sys, subprocess
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples')
for k, r in enumerate(rows[2:]):
    if len(r) < 10 or r[0] == 'Kernel Name':
        break
    print(f"{k:4d} ex={int(r[iex])/1e6:7.3f}M smp={r[ismp]:>5s}  {r[isrc].strip()[:100]}")
