"""SASS mnemonic counts of the production kernels in the built library (static, no GPU needed):
which kernels carry TMA bulk copies (UBLKCP: cp.async.bulk, both directions), TMA tensor loads (UTMALDG),
mbarrier traffic (SYNCS), the data-dependent slot release (REDUX), 256-bit global accesses (…ENL2.256) and how
much fp64 math (DFMA).  Usage: python scripts/sass_evidence.py > profiles/r1_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stefanproblem_b200", "lib", "libstefan_b200.so")
KEYS = ["UBLKCP", "UTMALDG", "SYNCS", "REDUX", "STG.E.ENL2.256", "LDG.E.ENL2.256", "DFMA", "SHFL", "LDS", "STS", "LDC", "BAR"]
WANT = ("apply_march", "fd_apply_vec", "fixup_tail", "blocks_apply", "blocks_assemble")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
op = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
counts, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = op.match(line)
    if m and fn:
        counts[fn][m.group(1)] += 1
names = list(counts)
demangled = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("kernel | SASS instructions | " + " | ".join(KEYS))
for n, d in sorted(zip(names, demangled), key=lambda t: t[1]):
    if not any(w in d for w in WANT):
        continue
    c = counts[n]
    short = re.sub(r"\(.*", "", d).replace("void stefan::", "")
    print(f"{short} | {sum(c.values())} | " + " | ".join(str(sum(v for o, v in c.items() if o.startswith(k))) for k in KEYS))
