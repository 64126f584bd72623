#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + hottest source lines by stall samples."""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
print("== %s" % rep)
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.02):
        print("%s [%s] = %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
while rows and (len(rows[0]) < 5 or rows[0][0] != "Address"):
    rows.pop(0)
if len(rows) > 2:
    hdr = rows[0]
    def col(name):
        for i, h in enumerate(hdr):
            if h.strip() == name:
                return i
        return None
    ci, cs, cx = col("Source"), col("Warp Stall Sampling (All Samples)"), col("Instructions Executed")
    if cs is None:
        cs = col("Warp Stall Sampling (All Cycles)")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = 0
    items = []
    for r in rows[1:]:
        try:
            smp = float(r[cs]); ex = float(r[cx]) if cx is not None else 0
        except Exception:
            continue
        tot += smp
        why = sorted(((float(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        items.append((smp, ex, r[ci].strip(), why, len(items)))
    opc = collections.Counter()
    opx = collections.Counter()
    for smp, ex, ins, why, idx in items:
        op = ins.split()[0] if ins.split() else "?"
        if op.startswith("@"):
            op = ins.split()[1]
        opc[op.split(".")[0]] += smp
        opx[op.split(".")[0]] += ex
    print("total samples", tot)
    print("-- samples by opcode:", [(k, round(100 * v / tot, 1)) for k, v in opc.most_common(14)])
    totx = sum(opx.values())
    print("-- executed by opcode:", [(k, round(100 * v / totx, 1)) for k, v in opx.most_common(14)])
    items.sort(reverse=True)
    for smp, ex, ins, why, idx in items[:topn]:
        print("%6.2f%%  #%-5d ex=%.3g  %-70s %s" % (100 * smp / tot, idx, ex, ins[:70], " ".join("%s=%d" % (w, v) for v, w in why if v)))
    # samples per block of 250 instructions (to attribute time to code regions)
    if len(sys.argv) > 3:
        B = int(sys.argv[3])
        by = collections.Counter()
        for smp, ex, ins, why, idx in items:
            by[idx // B] += smp
        print("-- samples per %d-instruction block:" % B)
        print(" ".join("%d:%.1f" % (k * B, 100 * v / tot) for k, v in sorted(by.items()) if v / tot > 0.002))
