"""Static look at a kernel's SASS: opcode histogram of the biggest loop (largest backward branch span).
usage: sass_loop.py lib.so 'regex of mangled kernel name'"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, ker = None, {}
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m: cur = m.group(1); ker[cur] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m and cur: ker[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in ker.items():
    if not re.search(pat, name): continue
    best = None
    for addr, t in ins:
        m = re.search(r'BRA(?:\.\S+)?\s+(?:\S+,\s+)?0x([0-9a-f]+)', t)
        if m and 'BRA.U.ANY' not in t:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]): best = (tgt, addr)
    print(name[:70], 'total', len(ins), 'loop', best and [hex(b) for b in best])
    if best:
        body = [t for a, t in ins if best[0] <= a <= best[1]]
        h = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for t in body)
        print('  loop instructions', len(body), dict(h.most_common(16)))
