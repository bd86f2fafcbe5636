"""Opcode histogram (by executed warp instructions) and top stall lines of an ncu source page.
usage: ncu_source.py report.ncu-rep [N]"""
import csv, subprocess, sys, collections, re
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
hist, samp = collections.Counter(), collections.Counter()
tot = 0
data = []
for r in rows[1:]:
    if len(r) <= iE: continue
    try: n = int(r[iE]); s = int(r[iN])
    except ValueError: continue
    src = r[iS].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2) if m else src
    base = op.split('.')[0]
    hist[base] += n; samp[base] += s; tot += n
    data.append((s, n, src, r))
print('total warp instructions', tot)
for k, v in hist.most_common(30): print(f'  {k:12s} {v:12d} {100*v/tot:5.1f}%  samples {samp[k]}')
print('top stall lines:')
for s, n, src, r in sorted(data, key=lambda t: -t[0])[:topn]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:3]
    print(f'  {s:6d} {n:10d} {src[:70]:70s} ', ' '.join(f'{h[6:]}={v}' for v, h in st if v))
