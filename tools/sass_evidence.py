"""cuobjdump -sass of libhandposedd.so -> per-kernel counts of the Blackwell-specific mnemonics (profiles/r1_sass_evidence.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hand_tracking_samples_b200", "libhandposedd.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, name = {}, None
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        kern[name][m.group(1)] += 1


def dem(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return (r or n).split("(")[0].replace("void hp::", "").replace("hp::", "")


print("# SASS evidence (cuobjdump -sass libhandposedd.so, sm_100a only)\n")
print("Counts of the Blackwell-specific mnemonics per kernel (B200_PROFILING.md: `tcgen05.mma` -> `UTCHMMA`, `tcgen05.ld` -> `LDTM`,")
print("TMA -> `UTMALDG`, `cp.async.bulk` -> `UBLKCP`); no `HMMA`/`HGMMA` anywhere: %s.  Checked by `tests/test_sass_evidence.py`;" % (
    "confirmed" if not any(c["HMMA"] or c["HGMMA"] for c in kern.values()) else "VIOLATED"))
print("regenerate with `python tools/sass_evidence.py > profiles/r1_sass_evidence.md`.\n")
print("| kernel | UTCHMMA | LDTM | UTMALDG | UBLKCP | UTCBAR | instructions |")
print("|---|---|---|---|---|---|---|")
for k, c in sorted(kern.items(), key=lambda kv: dem(kv[0])):
    if c["UTCHMMA"] or "peer_" in k:
        print("| `%s` | %d | %d | %d | %d | %d | %d |" % (dem(k), c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UBLKCP"], c["UTCBAR"], sum(c.values())))
print("\n%d kernels in the library in total." % len(kern))
