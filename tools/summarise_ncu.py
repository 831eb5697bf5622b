#!/usr/bin/env python
"""`ncu -i x.ncu-rep --page raw --csv` -> the handful of metrics DESIGN.md / profiles cite, one block per launch."""
import csv
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active", "sm__mem_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled"]
with open(sys.argv[1], newline="") as f:
    rd = list(csv.reader(f))
hdr = next(i for i, r in enumerate(rd) if r and r[0] == "ID")
names, units = rd[hdr], rd[hdr + 1]
for r in rd[hdr + 2:]:
    if len(r) < len(names):
        continue
    d = dict(zip(names, r))
    print(f"## {d.get('Kernel Name', '?')[:90]}  (id {d.get('ID')})")
    for n, u in zip(names, units):
        if any(n.startswith(k) or k in n for k in KEYS):
            print(f"  {n} = {d[n]} {u}")
