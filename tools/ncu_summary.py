"""Condenses an .ncu-rep (ncu --set full) into the per-kernel summary CSV kept under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum"]
cols = []
for w in WANT:
    m = [i for i, h in enumerate(hdr) if h == w or h.endswith("." + w)]
    if m: cols.append(m[0])
with open(out, "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow([hdr[i].split("TriageCompute.")[-1] for i in cols])
    wr.writerow([units[i] for i in cols])
    for r in body:
        wr.writerow([r[i] for i in cols])
print(open(out).read()[:3000])
