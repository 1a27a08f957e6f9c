#!/usr/bin/env python
"""Per-kernel histogram of the SASS opcodes that prove (or would disprove) a Blackwell-native kernel
(B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies,
UBLKCP = cp.async.bulk, HMMA/IMMA = legacy mma.sync, plus the instruction count.  Writes profiles/sass_opcodes.txt.

    python tools/sass_histogram.py [path/to/libtranshuman_b200.so]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "transhuman_b200", "libtranshuman_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA",
        "IMMA", "HGMMA", "FFMA2", "F2FP", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["_n"] += 1
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    cur[k + ("" if k not in ("UTCHMMA",) else (".2CTA" if ".2CTA" in m.group(2) else ""))] += 1
    out = [f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; tools/sass_histogram.py)",
           "# UTC*MMA = tcgen05.mma | LDTM = tcgen05.ld | UTMALDG = cp.async.bulk.tensor | UBLKCP = cp.async.bulk | "
           "HMMA/IMMA/HGMMA = legacy tensor paths (expected: none)", ""]
    for name, c in per.items():
        short = re.sub(r"\(.*", "", demangle(name))
        tags = "  ".join(f"{k}={v}" for k, v in c.items() if k != "_n")
        out.append(f"{short:<60s} instr={c['_n']:<6d} {tags}")
    legacy = sum(c[k] for c in per.values() for k in ("HMMA", "IMMA", "HGMMA"))
    out.append("")
    out.append(f"legacy tensor instructions (HMMA / IMMA / HGMMA) in the whole library: {legacy}")
    path = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
