#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libsaf_b200.so, how often the Blackwell-specific instructions occur
(packed fp32 FFMA2/FMUL2, bulk-TMA UBLKCP, tensor-map TMA UTMALDG, tcgen05 UTCHMMA / LDTM / UTCBAR, mbarrier SYNCS,
USETMAXREG) and the first occurrence of each.  usage: tools/sass_excerpts.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "spatially_aware_ai_b200", "libsaf_b200.so")
WANT = ["FFMA2", "FMUL2", "FADD2", "UBLKCP", "UTMALDG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS",
        "USETMAXREG", "REDG", "RED.", "ATOMG", "MUFU.RCP", "LDG.E.128", "STG.E.128", "LDS.128", "STL", "LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kern, counts, first, total = None, {}, {}, collections.Counter()
    for line in out.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kern = re.sub(r"\(.*", "", kern)
            counts[kern], first[kern] = collections.Counter(), {}
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if not m or kern is None:
            continue
        total[kern] += 1
        text = m.group(2)
        for w in WANT:
            if w in text:
                counts[kern][w] += 1
                first[kern].setdefault(w, "/*%s*/ %s" % (m.group(1), text.strip()))
    lines = ["SASS of %s (cuobjdump -sass, sm_100a): occurrences per kernel and the first instance of each" % os.path.basename(SO), ""]
    for k in sorted(counts, key=lambda k: -total[k]):
        if not counts[k]:
            continue
        lines.append("%s   [%d instructions]" % (k, total[k]))
        lines.append("    " + "  ".join("%s x%d" % (w, c) for w, c in counts[k].most_common()))
        for w in ("FFMA2", "FMUL2", "UBLKCP", "UTMALDG", "UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "USETMAXREG"):
            if w in first[k]:
                lines.append("      %s" % first[k][w][:150])
        lines.append("")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text[:3000])


if __name__ == "__main__":
    main()
