"""Per-source-line hot spots of one kernel from an .ncu-rep (no GPU needed).

ncu's CSV export of the source page carries metrics only in the SASS view, so this joins it with
`nvdisasm -g` line info of the same kernel taken from the built library (instruction order is identical).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep attloc_fwd [launch-skip] [--top 40]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "robust_e2e_gan_b200", "libre2e_b200.so")


def sass_rows(rep, pattern, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pattern,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hdr = rows[1]
    seen, uniq = set(), []
    for r in rows[2:]:                      # the export lists every instruction twice: keep the first
        if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            uniq.append(r)
    return name, hdr, uniq


def line_table(kernel_name):
    """instruction index -> (file, line, sass) of the kernel whose demangled name matches."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    # match on the template-free function name, then on the instruction count later
    base = re.search(r"(\w+)\s*(<\(|<[^u]|\()", kernel_name.replace("<unnamed>", "")).group(1)
    base = re.findall(r"(\w+)(?=<|\()", kernel_name.replace("<unnamed>", ""))[0] if base in ("void",) else base
    tables = []
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur, fn, loc = None, None, ("?", 0)
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
            if m:
                fn = m.group(1)
                cur = [] if base in fn else None
                if cur is not None:
                    tables.append((fn, cur))
                continue
            if cur is None:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                cur.append((loc[0], loc[1], m.group(2)))
    return tables


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    name, hdr, rows = sass_rows(rep, pattern, skip)
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tables = [t for t in line_table(name) if len(t[1]) == len(rows)]
    if not tables:
        print("no disassembled function with %d instructions matches %s" % (len(rows), name))
        return
    # template instantiations of equal length: pick by template args in the mangled name if possible
    args = re.findall(r"\(int\)(\d+)", name)
    pick = tables[0]
    for fn, t in tables:
        if all(("Li%sE" % a) in fn for a in args):
            pick = (fn, t)
            break
    table = pick[1]
    agg = collections.OrderedDict()
    for r, (f, l, s) in zip(rows, table):
        a = agg.setdefault((f, l), [0, 0, collections.Counter()])
        a[0] += int(r[si] or 0)
        a[1] += int(r[ii] or 0)
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                a[2][hdr[c][6:]] += v
    tot_s = sum(a[0] for a in agg.values()) or 1
    tot_i = sum(a[1] for a in agg.values()) or 1
    print("%s\n  samples %d, warp instructions %d" % (name, tot_s, tot_i))
    src = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src:
            p = os.path.join(ROOT, "robust_e2e_gan_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ""
        st = ",".join("%s:%d" % kv for kv in a[2].most_common(3))
        print("%5.1f%% smp %5.1f%% ins  %s:%-4d %-60s %s" % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, f, l,
                                                            text[:60], st))


if __name__ == "__main__":
    main()
