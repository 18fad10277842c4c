"""Summarise an ncu --set full report (raw CSV page) per kernel: the metrics quoted in DESIGN.md / profiles/."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
stall = [h for h in hdr if "stalled" in h and "per_issue_active" in h]
seen = set()
for d in data:
    key = d[idx["Kernel Name"]]
    if key in seen and "--all" not in sys.argv:
        continue
    seen.add(key)
    print("----")
    for w in want:
        if w in idx:
            print(f"  {w:72s} {d[idx[w]][:90]} {units[idx[w]]}")
    st = sorted(((float(d[idx[h]].replace(',', '')), h) for h in stall if d[idx[h]] not in ("", "n/a")), reverse=True)[:5]
    for v, h in st:
        print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:40s} {v:.2f}")
