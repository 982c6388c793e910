#!/usr/bin/env python
"""Opcode evidence from the built library: per kernel, how many tcgen05 MMAs (UTCHMMA), TMEM loads (LDTM), bulk async
copies (UBLKCP), tcgen05 commits (UTCBAR), mbarrier operations (SYNCS) and packed fp32 operations (FFMA2 / FADD2 /
FMUL2) the SASS holds.  usage: sass_histogram.py [libmscope_b200.so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "MUFU", "LDS", "ELECT"]


def histogram(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    per[cur][o] += 1
    return per


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "common_b200", "csrc", "libmscope_b200.so")
    per = histogram(path)
    tot = collections.Counter()
    print("%-72s %s" % ("kernel", " ".join("%7s" % o for o in OPS)))
    for k, c in per.items():
        if sum(c.values()) == 0:
            continue
        tot.update(c)
        print("%-72s %s" % (k[:72], " ".join("%7d" % c[o] for o in OPS)))
    print("%-72s %s" % ("TOTAL", " ".join("%7d" % tot[o] for o in OPS)))
