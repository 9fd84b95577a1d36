#!/bin/bash
# Condenses the round-2 ncu captures in gpurun_out/ (tools/prof_r02.sh) into profiles/r02_*.txt.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
sum() { python tools/ncu_summary.py $G/$1.ncu-rep $2 $3 > $P/$1_summary.txt; }
sum r02_se99_n32_fast 131072 32
python tools/ncu_hot.py $G/r02_se99_n32_fast.ncu-rep 131072 40 > $P/r02_se99_n32_fast_instmix.txt
python tools/ncu_lines.py $G/r02_se99_n32_fast.ncu-rep 131072 80 > $P/r02_se99_n32_fast_lines.txt
python tools/ncu_smem.py $G/r02_se99_n32_fast.ncu-rep 131072 20 > $P/r02_se99_n32_fast_smem.txt
sum r02_se99_n32_fast_wishart 131072 32
sum r02_se99_n32_strict 131072 32
sum r02_verify_n32 131072 32
sum r02_gmw81_n256_fast 1184 256
BY_STALL=1 python tools/ncu_lines.py $G/r02_gmw81_n256_fast.ncu-rep 1184 30 > $P/r02_gmw81_n256_fast_lines.txt
sum r02_se99_n256_fast 1184 256
sum r02_se90_n64_fast 65536 64
sum r02_se99_n16_fast 262144 16
[ -f $G/r02_lwm1.ncu-rep ] && python tools/ncu_summary.py $G/r02_lwm1.ncu-rep 1184 256 > $P/r02_gmw81_n256_fast_warp_panel_summary.txt
python - <<'PY'
import csv, collections, json, re
rows = list(csv.reader(open("gpurun_out/r02_launches.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) != len(h): continue
    d = dict(zip(h, r))
    if d.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])[:90]
    v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open("profiles/r02_bench_launch_list.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs`\n")
    f.write(f"{'launches':>8s} {'total ms':>10s} {'share':>7s}  kernel\n")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{c:8d} {t:10.3f} {100 * t / tot:6.1f}%  {k}\n")
PY
python - <<'PY'
import subprocess, csv, io, json
raw = subprocess.run(["ncu", "-i", "gpurun_out/r02_se99_n32_fast.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
def b(k):
    v = float(d[k].replace(",", "")); return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u[k], 1)
rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
json.dump({"source": "ncu --set full, profiles/r02_se99_n32_fast_summary.txt (wip_kernel<SE99, n=32>, 131072 matrices)",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "matrices": 131072,
           "dram_bytes_per_matrix": (rd + wr) / 131072, "algorithmic_bytes_per_matrix": 16896,
           "note": "reads are below the algorithmic 8 n^2 because only the lower triangle of the symmetric input is fetched"},
          open("profiles/r02_traffic.json", "w"), indent=1)
PY
cp $G/r02_sweep_a.txt $P/r02_size_sweep.txt
[ -f $G/r02_sweep_warp_panel.txt ] && cp $G/r02_sweep_warp_panel.txt $P/r02_size_sweep_warp_panel.txt
ls $P | grep r02
