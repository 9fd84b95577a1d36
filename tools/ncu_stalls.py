"""Warp-stall samples per CUDA source line, and per SASS instruction for the hottest lines.
usage: python tools/ncu_stalls.py report.ncu-rep [top]"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
out = []
for rec in rows:
    if rec and rec[0] in ("Address", "#"):
        hdr = rec
        continue
    if hdr is None or len(rec) != len(hdr):
        continue
    d = dict(zip(hdr, rec))
    try:
        st = int(d.get("Warp Stall Sampling (All Samples)", "0") or 0)
        ex = int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        continue
    out.append((st, ex, d.get("Address", ""), d.get("Source", "")))
tot = sum(o[0] for o in out) or 1
print(f"total samples {tot}")
for i, (st, ex, addr, src) in enumerate(out):
    pass
idx = sorted(range(len(out)), key=lambda i: -out[i][0])[:top]
for i in idx:
    st, ex, addr, src = out[i]
    prev = out[i - 1][3][:60] if i else ""
    print(f"{100 * st / tot:5.1f}%  ex={ex:9d}  {src[:70]:70s}  <- {prev}")
