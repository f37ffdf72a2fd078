"""Per-source-line summary of an `ncu --page source --csv --print-source sass,cuda` dump: instructions executed and
stall samples aggregated per CUDA line (top N).  usage: ncu_src_summary.py dump.csv [file-substring] [N]"""
import csv, sys
path = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if r[0] and hdr:           # a CUDA source line (aggregated)
        try:
            inst, samp = int(r[i_inst]), int(r[i_samp])
        except ValueError:
            continue
        stalls = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:3]
        out.append((samp, inst, cur_file.split("/")[-1], r[0], r[1].strip()[:90], stalls))
tot_s = sum(o[0] for o in out) or 1
tot_i = sum(o[1] for o in out) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
for o in sorted(out, reverse=True):
    if only and only not in o[2]:
        continue
    if topn <= 0:
        break
    topn -= 1
    print("%5.1f%% samp %5.1f%% inst  %s:%s  %s   %s" % (100.0 * o[0] / tot_s, 100.0 * o[1] / tot_i, o[2], o[3], o[4],
                                                      " ".join("%s=%d" % (h[6:], v) for v, h in o[5] if v)))
