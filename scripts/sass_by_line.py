"""Join an ncu SASS source-page CSV with nvdisasm line info: executed warp-instructions per source line.
usage: sass_by_line.py <ncu_source.csv> <cubin> <mangled kernel substring> <agents>"""
import csv, re, subprocess, sys, collections
csv_path, cubin, kname, agents = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(csv_path)))
secs = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'rows': []}; secs.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and len(r) > 5: cur['rows'].append(r)
s = secs[0]; h = s['hdr']; ie = h.index('Instructions Executed'); src = h.index('Source')
counts = [int(r[ie]) for r in s['rows']]
txt = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout.splitlines()
# find the function's text section
start = None
for i, l in enumerate(txt):
    if l.startswith('.text.') and kname in l: start = i
    elif l.strip().startswith('.section') and '.text.' in l and kname in l: start = i
assert start is not None
lines = []; curline = ('?', 0)
for l in txt[start + 1:]:
    if l.strip().startswith('.section') or (l.startswith('.text.') and kname not in l):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curline = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): lines.append(curline)
print('sass rows', len(counts), 'disasm instrs', len(lines))
n = min(len(counts), len(lines))
W = agents / 32
agg = collections.Counter()
for c, ln in zip(counts[:n], lines[:n]): agg[ln] += c
tot = sum(agg.values())
srcs = {}
for (f, ln), c in sorted(agg.items(), key=lambda kv: -kv[1])[:45]:
    if f not in srcs:
        try: srcs[f] = open('/root/repo/gradabm-june_b200/csrc/' + f).read().splitlines()
        except Exception: srcs[f] = []
    code = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ''
    print(f'{c / W:8.1f} {100 * c / tot:5.1f}%  {f}:{ln:<5d} {code}')
