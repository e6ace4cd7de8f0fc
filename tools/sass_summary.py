#!/usr/bin/env python
"""Blackwell-native evidence without rebuilding: disassembles ccv_b200/libccv_nnc_sm100.so (cuobjdump -sass) and counts, per
kernel, the SASS mnemonics that B200_PROFILING.md names -- UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG /
UTMASTG (TMA loads / stores, incl. the .IM2COL forms), UBLKCP, UTCBAR (tcgen05.commit), plus HMMA (legacy mma.sync: must be 0).
Writes a table (one line per kernel that uses any of them, and the totals) to stdout:  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ccv_b200", "libccv_nnc_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
pat = collections.OrderedDict([("UTC*MMA", re.compile(r"\bUTC[A-Z]*MMA")), ("LDTM", re.compile(r"\bLDTM")), ("STTM", re.compile(r"\bSTTM")), ("UTMALDG", re.compile(r"\bUTMALDG")),
                               ("UTMALDG.IM2COL", re.compile(r"\bUTMALDG\S*IM2COL")), ("UTMASTG", re.compile(r"\bUTMASTG")), ("UBLKCP", re.compile(r"\bUBLKCP")), ("UTCBAR", re.compile(r"\bUTCBAR")),
                               ("HMMA", re.compile(r"\bHMMA")), ("HGMMA", re.compile(r"\b[HQI]GMMA"))])
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pat.items():
            if p.search(line):
                kernels[cur][k] += 1
demangled = {}
try:
    names = list(kernels)
    dm = subprocess.run(["c++filt"] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
    demangled = dict(zip(names, dm))
except Exception:
    pass
tot = collections.Counter()
print("SASS mnemonic counts of %s (sm_100a), %d kernels in the image" % (os.path.basename(so), len(kernels)))
print(" ".join("%16s" % k for k in pat) + "  kernel")
for name, c in kernels.items():
    tot.update(c)
    if sum(c.values()):
        d = demangled.get(name, name)
        print(" ".join("%16d" % c[k] for k in pat) + "  " + (d[:150] + "..." if len(d) > 150 else d))
print(" ".join("%16d" % tot[k] for k in pat) + "  TOTAL")
