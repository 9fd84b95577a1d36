"""Shared-memory wavefronts per CUDA source line (ideal and excessive) from an .ncu-rep captured
with --import-source on.  usage: python tools/ncu_smem.py report.ncu-rep batch [top]"""
import csv
import io
import os
import subprocess
import sys

rep, batch = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows, fname, hdr = [], "?", None
for rec in csv.reader(io.StringIO(raw)):
    if not rec:
        continue
    if rec[0] == "File Path":
        fname = os.path.basename(rec[1])
        continue
    if rec[0] == "Line No":
        hdr = rec
        continue
    if hdr is None or not rec[0].isdigit():
        continue
    def num(name):
        v = rec[hdr.index(name)]
        return int(v) if v.isdigit() else 0
    w, ex, ideal = num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Excessive"), num("L1 Wavefronts Shared Ideal")
    if w:
        rows.append((w, ex, ideal, num("Instructions Executed"), fname, int(rec[0]), rec[1].strip()))
tot, totx = sum(r[0] for r in rows), sum(r[1] for r in rows)
print(f"shared wavefronts per matrix {tot / batch:.0f}, excessive {totx / batch:.0f}")
print(f"{'wavefronts':>10s} {'excess':>8s} {'ideal':>8s} {'instr':>7s}  location")
for w, ex, ideal, ins, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{w / batch:10.1f} {ex / batch:8.1f} {ideal / batch:8.1f} {ins / batch:7.1f}  {f}:{ln}  {src[:80]}")
