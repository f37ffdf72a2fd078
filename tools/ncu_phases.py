"""Instructions / stall samples of one kernel grouped by source-line ranges ("phases").

    python tools/ncu_phases.py REP KERNEL_REGEX SKIP FILE.cu "label=substring of first line" ...
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as n  # noqa: E402


def main():
    rep, pat, skip, fname = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    marks_in = [a.split("=", 1) for a in sys.argv[5:]]
    name, hdr, rows = n.sass_rows(rep, pat, skip)
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    tabs = [t for t in n.line_table(name) if len(t[1]) == len(rows)]
    import re
    args = re.findall(r"\(int\)(\d+)", name)
    pick = tabs[0]
    for fn, t in tabs:
        if "".join("Li%sE" % a for a in args) in fn:
            pick = (fn, t)
            break
    src = open(os.path.join(n.ROOT, "robust_e2e_gan_b200", "csrc", fname)).read().splitlines()
    start = 0
    for i, l in enumerate(src):
        if re.search(r"(\w+_kernel)", name).group(1) + "(" in l and "__global__" in "".join(src[max(0, i - 2):i + 1]):
            start = i
            break
    marks = []
    for lab, sub in marks_in:
        for i in range(start, len(src)):
            if sub in src[i]:
                marks.append((lab, i + 1))
                break
    ph, phs = collections.OrderedDict(), collections.Counter()
    last = 0
    for r, (f, l, s) in zip(rows, pick[1]):
        if f == fname:
            last = l
        lab = "(before)"
        for nm, ln in marks:
            if last >= ln:
                lab = nm
        ph[lab] = ph.get(lab, 0) + int(r[ii] or 0)
        phs[lab] += int(r[si] or 0)
    tot, tots = sum(ph.values()), sum(phs.values())
    print(name, "instr", tot, "samples", tots)
    for k in ph:
        print("%-16s %5.1f%% instr  %5.1f%% samples" % (k, 100.0 * ph[k] / tot, 100.0 * phs[k] / tots))


if __name__ == "__main__":
    main()
