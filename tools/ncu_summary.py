"""Condense an .ncu-rep into the handful of numbers the round notes quote.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [batch] [n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else None
n = int(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_branch.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:100])
    for k in KEYS:
        if k in d:
            print(f"  {k:72s} {d[k]:>16s} {units[hdr.index(k)]}")
    stalls = [(float(v.replace(',', '')), h) for h, v in d.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v]
    print("  stall cycles per issued instruction (top):")
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
    if batch:
        inst = float(d["smsp__inst_executed.sum"].replace(',', ''))
        t = float(d["gpu__time_duration.sum"].replace(',', ''))
        tu = units[hdr.index("gpu__time_duration.sum")]
        ts = t * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1, "nsecond": 1e-9}.get(tu, 1e-6)
        print(f"  warp-instructions per matrix: {inst / batch:.0f};  matrices/s under ncu: {batch / ts / 1e6:.2f} M")
        if n:
            alg = batch * (16 * n * n + 16 * n)
            rd = d.get("dram__bytes_read.sum", "0"); wr = d.get("dram__bytes_write.sum", "0")
            print(f"  algorithmic bytes {alg/1e9:.3f} GB; dram read {rd} + write {wr} {units[hdr.index('dram__bytes_read.sum')]}")
