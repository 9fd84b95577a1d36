"""Dynamic instruction mix and hot regions from `ncu --page source --csv`.
usage: python tools/ncu_hot.py report.ncu-rep batch [top]"""
import csv
import io
import re
import subprocess
import sys

rep, batch = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
ops = {}
tot = 0
stall = {}
for r in rows:
    ex = int(r["Instructions Executed"] or 0)
    src = r["Source"].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    ops[op] = ops.get(op, 0) + ex
    tot += ex
    stall[op] = stall.get(op, 0) + int(r["Warp Stall Sampling (All Samples)"] or 0)
print(f"total warp-instr {tot}, per matrix {tot / batch:.0f}")
st = sum(stall.values())
print(f"{'op':14s} {'instr/matrix':>12s} {'%':>6s} {'stall%':>7s}")
for op, c in sorted(ops.items(), key=lambda x: -x[1])[:top]:
    print(f"{op:14s} {c / batch:12.1f} {100 * c / tot:6.1f} {100 * stall[op] / max(st, 1):7.1f}")
