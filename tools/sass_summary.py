#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), FFMA2/FADD2/FMUL2 (packed fp32x2),
SYNCS (mbarrier), HMMA (legacy mma.sync, must be 0).  Runs on the build host:  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "jdet_b200", "_C", "libjdet_b200.so")
sass = subprocess.check_output(["cuobjdump", "-sass", so]).decode(errors="replace")
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|FFMA2|FADD2|FMUL2|SYNCS[.\w]*|HMMA[.\w]*|UTCBAR[.\w]*|REDG[.\w]*|ATOMG[.\w]*)")
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.check_output(["c++filt", m.group(1)]).decode().strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m:
        key = m.group(1)
        key = re.sub(r"^(SYNCS)\..*", r"\1", key)
        key = re.sub(r"^(REDG|ATOMG)\..*", r"\1", key)
        counts[cur][key] += 1
print("# SASS evidence per kernel of jdet_b200/_C/libjdet_b200.so (cuobjdump -sass; sm_100a)\n")
cols = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG.2D.GATHER4", "UBLKCP.S.G", "UBLKCP.G.S", "FFMA2", "FADD2", "FMUL2", "SYNCS", "REDG", "HMMA"]
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
for k, c in counts.items():
    merged = collections.Counter()
    for name, v in c.items():
        hit = [col for col in cols if name.startswith(col)]
        merged[hit[0] if hit else name] += v
    if not any(merged.values()):
        continue
    extra = {n: v for n, v in merged.items() if n not in cols}
    print("| `%s` | " % k[:90] + " | ".join(str(merged.get(col, 0)) for col in cols) + " |" + ("  <!-- %s -->" % extra if extra else ""))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nTotals:", dict(tot))
