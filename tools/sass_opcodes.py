#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05 MMA / TMEM
load-store, TMA tensor and bulk copies, bulk reduce, cluster barriers, distributed-shared-memory stores -- read from
`cuobjdump -sass` of the in-tree library.  Writes the table to stdout (committed as profiles/r02_sass_opcodes.txt).

    python tools/sass_opcodes.py [path/to/libre2e_b200.so]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "UCGABAR", "SYNCS", "STAS",
       "LDGSTS", "FFMA2", "MUFU"]


def kernel_opcodes(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    return counts


def demangle(names):
    try:
        r = subprocess.run(["cu++filt"] + names, capture_output=True, text=True)
        d = r.stdout.splitlines()
        if len(d) == len(names):
            return d
    except OSError:
        pass
    return names


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "robust_e2e_gan_b200", "libre2e_b200.so")
    counts = kernel_opcodes(lib)
    names = list(counts)
    short = []
    for n in demangle(names):
        n = re.sub(r"\(anonymous namespace\)::|re2e::|<unnamed>::", "", n)
        n = re.sub(r"\((int|bool|unsigned int)\)", "", n)
        n = re.sub(r"\(.*\)$", "", n).replace("void ", "")
        short.append(n[:64])
    print("SASS opcode counts per kernel of %s (cuobjdump -sass; sm_100a)" % os.path.basename(lib))
    print("%-64s " % "kernel" + " ".join("%8s" % o for o in OPS))
    tot = collections.Counter()
    for n, s in zip(names, short):
        c = counts[n]
        tot.update(c)
        if sum(c.values()):
            print("%-64s " % s + " ".join("%8d" % c[o] for o in OPS))
    print("%-64s " % "TOTAL" + " ".join("%8d" % tot[o] for o in OPS))


if __name__ == "__main__":
    main()
