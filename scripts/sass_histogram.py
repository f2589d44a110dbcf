"""Per-kernel SASS opcode histogram of the in-tree library (evidence for DESIGN.md section 3: which kernels use
UTCHMMA / UTMALDG / LDTM / STTM / UTCBAR, and that none uses the legacy HMMA path).
usage: python scripts/sass_histogram.py [lib.so] > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ctrl_adapter_b200", "libctrl_adapter_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "UTCBAR", "LDTM", "STTM", "SYNCS", "MUFU", "FFMA2", "HMMA", "LDGSTS", "BAR"]
kern, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            hist[kern]["UTCHMMA.2CTA"] += 1
print("# SASS opcode histogram per kernel (cuobjdump -sass of the in-tree libctrl_adapter_b200.so, sm_100a)\n")
print("Counts of the Blackwell-specific opcodes (tcgen05.mma = UTCHMMA, TMA = UTMALDG, cp.async.bulk 1-D = UBLKCP, tcgen05.ld/st = LDTM/STTM, "
      "tcgen05.commit = UTCBAR, mbarrier = SYNCS) and of the legacy tensor path (HMMA) per compiled kernel; `instrs` = all.\n")
print("| kernel | instrs | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
tot = collections.Counter()
for k, c in hist.items():
    n = sum(v for kk, v in c.items() if kk != "UTCHMMA.2CTA")
    print(f"| `{k[:90]}` | {n} | " + " | ".join(str(c.get(x, 0)) for x in KEY) + " |")
    tot.update(c)
print(f"| **total ({len(hist)} kernels)** | {sum(v for kk, v in tot.items() if kk != 'UTCHMMA.2CTA')} | " +
      " | ".join(str(tot.get(x, 0)) for x in KEY) + " |")
