#!/usr/bin/env python
"""Stall samples of an .ncu-rep aggregated by CUDA source line (needs -lineinfo and --import-source on)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = []
tot = 0
cur_file = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s, i_x, i_w, i_wi = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        smp, ex = float(r[i_s] or 0), float(r[i_x] or 0)
        wv, wvi = float(r[i_w] or 0), float(r[i_wi] or 0)
    except ValueError:
        continue
    agg.append((smp, ex, wv, wvi, cur_file, r[0], r[1].strip()[:110]))
    tot += smp
agg.sort(key=lambda t: -t[0])
print("total samples %.0f" % tot)
for smp, ex, wv, wvi, f, ln, src in agg[:topn]:
    print("%5.2f%%  ex=%.2e  smem_wf=%.2e (ideal %.2e)  %s:%s  %s" % (100 * smp / tot, ex, wv, wvi, f, ln, src))
