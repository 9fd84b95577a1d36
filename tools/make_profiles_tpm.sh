#!/bin/bash
# Condenses the final captures of tools/prof_final.sh (thread-per-matrix kernels, launch list,
# sweep, bench line) into profiles/.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
for f in r02_tpr_se99_n4:4194304 r02_tpr_se99_n8:1048576 r02_tps_se99_n12:1048576; do
  n=${f%%:*}; b=${f##*:}
  python tools/ncu_summary.py $G/$n.ncu-rep $b > $P/${n}_summary.txt
done
python tools/ncu_hot.py $G/r02_tpr_se99_n4.ncu-rep 4194304 25 > $P/r02_tpr_se99_n4_instmix.txt
python tools/ncu_hot.py $G/r02_tps_se99_n12.ncu-rep 1048576 25 > $P/r02_tps_se99_n12_instmix.txt
python - <<'PY'
import csv, collections, re
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
cp $G/r02_sweep_a.txt $P/r02_size_sweep.txt
cp $G/r02_bench_line.json $P/r02_bench_line.json
