#!/usr/bin/env python
"""SASS evidence per kernel of the in-tree library (cuobjdump -sass, no GPU needed): instruction counts of the mnemonics that prove
the Blackwell paths (UTCHMMA = tcgen05.mma, LDTM / STTM = TMEM access, UTCBAR = tcgen05.commit, UBLKCP = TMA bulk copy, UBLKRED =
TMA bulk reduction, ATOMS = shared-memory atomics, REDG = global reductions, UCGABAR = cluster barrier, SYNCS = mbarrier) plus the
register / shared-memory footprint from the ptxas logs.  usage: python tools/sass_evidence.py > profiles/<tag>_sass_evidence.txt"""
import re, subprocess, sys
from collections import Counter, OrderedDict
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "ro_map_b200" / "libmon_b200.so"
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UBLKRED", "UTMALDG", "ATOMS", "REDG", "RED.", "ATOMG", "UCGABAR", "SYNCS",
             "LDGSTS", "HMMA", "LDS", "STS", "LDG", "STG", "F2I", "I2F", "MUFU", "BAR.SYNC", "ACQBULK", "DEPBAR", "ERRBAR"]
def short(name):
    """kernel name with its template arguments, without the parameter list"""
    name = name.replace("void ", "")
    m = re.match(r"(\w+<.*?>)\(", name)
    return (m.group(1) if m else re.sub(r"\(.*", "", name)).replace("(int)", "").replace("(bool)", "")


sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
kern = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                kern[cur][mn] += 1
demangle = subprocess.run(["cu++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines() if kern else []
names = dict(zip(kern, demangle)) if len(demangle) == len(kern) else {k: k for k in kern}
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# {LIB.relative_to(ROOT)}: {len(kern)} kernels, cubin architectures {arch}")
tot = Counter()
for k, c in kern.items():
    tot.update(c)
    n = short(names[k])
    hits = ", ".join(f"{mn} x{c[mn]}" for mn in MNEMONICS if c[mn])
    print(f"{n:44s} {c['_total']:6d} SASS instructions: {hits}")
print("# whole library: " + ", ".join(f"{mn} x{tot[mn]}" for mn in MNEMONICS if tot[mn]))
print("# registers / shared memory / spills (ptxas -v, ro_map_b200/_build/*.ptxas.log)")
for log in sorted((ROOT / "ro_map_b200" / "_build").glob("*.ptxas.log")):
    txt = log.read_text().splitlines()
    for i, l in enumerate(txt):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", l)
        if m:
            used = next((x for x in txt[i:i + 6] if "Used" in x), "")
            spill = next((x for x in txt[i:i + 6] if "spill" in x), "")
            nm = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            print(f"{short(nm):44s} {used.split(':', 1)[-1].strip()}; {spill.strip()}")
