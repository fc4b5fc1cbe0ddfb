#!/usr/bin/env python
"""Opcode histogram per kernel of a built library (cuobjdump -sass), for profiles/: which kernels carry the tensor-pipe
(DMMA / UTCIMMA), TMEM (LDTM), TMA bulk-copy (UBLKCP), cp.async (LDGSTS) and mbarrier (SYNCS) instructions."""
import collections
import re
import subprocess
import sys

WATCH = ["UTCIMMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "DMMA", "HMMA", "IMMA", "LDGSTS", "SYNCS", "DFMA", "DMUL",
         "DADD", "MUFU", "I2F", "F2I", "BAR", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "SHFL"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    print("library: %s" % path)
    print("%-78s %7s  %s" % ("kernel", "instr", "watched opcodes"))
    tot = collections.Counter()
    for (fn, h), name in zip(hist.items(), demangle):
        n = sum(h.values())
        short = re.sub(r"\(.*", "", name)[:78]
        items = ["%s %d" % (k, h[k]) for k in WATCH if h.get(k)]
        print("%-78s %7d  %s" % (short, n, ", ".join(items)))
        tot.update(h)
    print("\nwhole library: " + ", ".join("%s %d" % (k, tot[k]) for k in WATCH if tot.get(k)))


if __name__ == "__main__":
    main(sys.argv[1])
