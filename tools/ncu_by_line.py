#!/usr/bin/env python3
"""Top source lines of an ncu report by warp-stall samples (cuda,sass source view).
usage: tools/ncu_by_line.py report.ncu-rep [top] [file-substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fpath, hdr, rows = "", None, []
for rec in csv.reader(io.StringIO(out)):
    if len(rec) == 2 and rec[0] == "File Path":
        fpath = rec[1].split("/")[-1]
        continue
    if rec and rec[0] == "Line No":
        hdr = rec
        continue
    if hdr is None or len(rec) < len(hdr) or not rec[0].strip().isdigit():
        continue
    d = dict(zip(hdr[4:], rec[4:]))
    def g(k):
        try:
            return int(d.get(k) or 0)
        except ValueError:
            return 0
    rows.append((fpath, int(rec[0]), rec[1].strip()[:70], g("# Samples"), g("Instructions Executed"), g("stall_long_sb"),
                 g("stall_wait"), g("stall_barrier"), g("stall_no_inst"), g("L2 Theoretical Sectors Local"), g("L2 Theoretical Sectors Global")))
ts = sum(r[3] for r in rows) or 1
ti = sum(r[4] for r in rows) or 1
print("total samples %d, instructions %d" % (ts, ti))
print("%-18s %5s %6s %6s %5s %5s %5s %5s %9s %9s  %s" % ("file", "line", "samp%", "inst%", "lsb%", "wait%", "bar%", "noi%", "L2loc", "L2glob", "source"))
for r in sorted([r for r in rows if only in r[0]], key=lambda r: -r[3])[:top]:
    s = max(r[3], 1)
    print("%-18s %5d %5.2f%% %5.2f%% %4.0f%% %4.0f%% %4.0f%% %4.0f%% %9d %9d  %s" % (r[0][:18], r[1], 100.0 * r[3] / ts, 100.0 * r[4] / ti, 100.0 * r[5] / s,
          100.0 * r[6] / s, 100.0 * r[7] / s, 100.0 * r[8] / s, r[9], r[10], r[2]))
