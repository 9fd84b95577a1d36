"""Executed warp instructions per CUDA source line (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep batch [top]"""
import csv
import io
import os
import subprocess
import sys

rep, batch = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = []
fname = "?"
hdr = None
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
    ex = num("Instructions Executed")
    st = num("Warp Stall Sampling (All Samples)")
    if ex:
        rows.append((ex, st, fname, int(rec[0]), rec[1].strip()))
tot = sum(r[0] for r in rows)
stt = sum(r[1] for r in rows)
print(f"total warp-instr {tot}, per matrix {tot / batch:.0f}")
print(f"{'instr/matrix':>12s} {'%':>5s} {'stall%':>6s}  location")
for ex, st, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{ex / batch:12.1f} {100 * ex / tot:5.1f} {100 * st / max(stt, 1):6.1f}  {f}:{ln}  {src[:90]}")
if os.environ.get("BY_STALL"):
    print("by stall samples:")
    for ex, st, f, ln, src in sorted(rows, key=lambda r: -r[1])[:top]:
        print(f"{ex / batch:12.1f} {100 * ex / tot:5.1f} {100 * st / max(stt, 1):6.1f}  {f}:{ln}  {src[:90]}")
if len(sys.argv) > 4:   # region sums: file:lo-hi,...
    print("regions:")
    for spec in sys.argv[4].split(","):
        f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        ex = sum(r[0] for r in rows if r[2] == f and lo <= r[3] <= hi)
        stx = sum(r[1] for r in rows if r[2] == f and lo <= r[3] <= hi)
        print(f"  {spec:28s} {ex / batch:9.1f}  {100 * ex / tot:5.1f}%  stall {100 * stx / max(stt, 1):5.1f}%")
