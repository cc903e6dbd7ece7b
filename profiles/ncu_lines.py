"""Per-source-line instruction / sample counts of one kernel of an ncu report.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [min-fraction]
Joins the report's SASS page (Address, Instructions Executed, # Samples) with the line table that
`nvdisasm -g` prints for the kernel (compiled with -lineinfo)."""
import csv, re, subprocess, sys, tempfile, os, collections

rep, lib, kname = sys.argv[1:4]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's text section
line_of = {}
cur, inside = None, False
for l in dis:
    if l.startswith('.text.') or l.lstrip().startswith('.section'):
        inside = kname in l and '.text.' in l
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and cur is not None:
        line_of[int(m.group(1), 16)] = cur
csrc = os.path.join(os.path.dirname(os.path.abspath(lib)), '..', 'csrc')
src_files = {f: open(os.path.join(csrc, f)).read().splitlines() for f in os.listdir(csrc) if f.endswith(('.cu', '.cuh'))}
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, iex, ismp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
agg = collections.defaultdict(lambda: [0, 0])
base = None
tot = 0
for r in rows[2:]:
    if len(r) <= iex or r[0] == 'Kernel Name' or r[0] == 'Address':
        break
    try:
        addr = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = addr
    ln = line_of.get(addr - base, ('?', -1))
    agg[ln][0] += int(r[iex]); agg[ln][1] += int(r[ismp]); tot += int(r[iex])
print(f'total warp instructions {tot/1e6:.2f}M')
for key in sorted(agg):
    ex, smp = agg[key]
    f, ln = key
    if ex >= tot * thr:
        text = src_files[f][ln - 1].strip()[:100] if f in src_files and 0 < ln <= len(src_files[f]) else f
        print(f'{ex/1e6:8.2f}M {100*ex/tot:5.1f}% smp={smp:6d}  L{ln}: {text}')
