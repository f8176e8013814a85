#!/usr/bin/env python
"""Which hardware paths each kernel of libmnv1.so uses, read off its SASS (cuobjdump, CPU only):
tcgen05.mma = UTCHMMA / UTCIMMA..., TMEM loads = LDTM, TMA = UTMALDG / UTMASTG, bulk copies = UBLKCP,
cp.async = LDGSTS, DP4A = IDP.4A, packed FP32 = FFMA2, 256-bit global stores = STG.E.ENL2.256, legacy mma.sync = HMMA.
usage: python tools/sass_mnemonics.py [libmnv1.so] > profiles/rNN_sass_mnemonics.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else glob.glob(os.path.join(ROOT, "cnn-*", "libmnv1.so"))[0]
KEYS = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "IDP.4A", "FFMA2", "HMMA", "ENL2.256", "SYNCS"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        counts[cur]["instr"] += 1
        for k in KEYS:
            if (" HMMA" in line) if k == "HMMA" else (k in line):     # HMMA (mma.sync) is not UTCHMMA
                counts[cur][k] += 1
print(f"# {os.path.basename(lib)}: SASS mnemonics per kernel (tools/sass_mnemonics.py); columns: static instruction counts")
def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("mnv1::", "").replace("void ", "")
    name = name.replace("(int)", "").replace("(bool)", "")
    depth = 0
    for i, ch in enumerate(name):            # the parameter list starts at the first '(' outside the template arguments
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            return name[:i]
    return name


seen = set()
print("%-100s %6s " % ("kernel", "instr") + " ".join("%8s" % k for k in KEYS))
for fn in order:
    if fn in seen:
        continue
    seen.add(fn)
    c = counts[fn]
    print("%-100s %6d " % (short(demangle(fn))[-100:], c["instr"]) + " ".join("%8s" % (c[k] or ".") for k in KEYS))
