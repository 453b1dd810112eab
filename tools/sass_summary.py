"""Counts the Blackwell-native SASS mnemonics per kernel of libdistrifuser_b200.so (cuobjdump -sass): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UTCBAR (tcgen05.commit), SYNCS (mbarrier), USETMAXREG (setmaxnreg), LDL / STL (spills), plus legacy HMMA (must be 0).
    python tools/sass_summary.py > profiles/r2_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "distrifuser_b200", "libdistrifuser_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "HGMMA", "MUFU", "FFMA2", "FADD2",
         "UCGABAR", "ACQBULK", "RED", "ATOM", "USETMAXREG", "FMNMX3", "LDL", "STL")
cur, per = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        while cur in per:
            cur += "'"
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        per[cur]["_total"] += 1
        for w in WATCH:
            if m.group(1).startswith(w):
                per[cur][w] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instruction counts per kernel (sm_100a)")
for k, c in per.items():
    tags = "  ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{k:60s} total={c['_total']:5d}  {tags}")
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# whole library: " + "  ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]) + f"  (legacy HMMA={tot['HMMA']}, HGMMA={tot['HGMMA']})")
