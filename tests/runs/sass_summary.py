"""Per-kernel SASS opcode summary of libsamq.so (tcgen05 / TMEM / TMA evidence):
    python tests/runs/sass_summary.py > profiles/rNN_sass_opcodes.txt
UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA
load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2/FMUL2/FADD2 = packed fp32."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sam_quantization_b200", "lib", "libsamq.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU", "FFMA2", "HFMA2", "total"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void samq::", "").replace("void ", "")
        name = re.sub(r"\(.*$", "", name)
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        base = op.split(".")[0]
        cur[base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
print(f"# SASS opcode counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)")
print(f"{'kernel':70s} " + " ".join(f"{k:>12s}" for k in KEYS))
tot = collections.Counter()
for name, c in per.items():
    print(f"{name[:70]:70s} " + " ".join(f"{c.get(k, 0):12d}" for k in KEYS))
    for k in KEYS:
        tot[k] += c.get(k, 0)
print(f"{'ALL':70s} " + " ".join(f"{tot[k]:12d}" for k in KEYS))
