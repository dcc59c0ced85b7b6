"""Turn an .ncu-rep (ncu --set full) into the markdown summary kept under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex] > profiles/xyz.md"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?")
    if pat and not pat.search(name):
        continue
    print(f"## {name.split('(')[0]}  (launch id {d.get('ID')})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
    st = [(float(d[h]), h) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and
          not h.endswith("not_issued") and d.get(h, "") not in ("", "n/a")]
    tot = sum(x for x, _ in st) or 1.0
    print("\nwarp stall samples (share of all samples):\n")
    for x, h in sorted(st, reverse=True)[:9]:
        print(f"- {h.replace('smsp__pcsamp_warps_issue_stalled_', '')}: {100 * x / tot:.1f}%")
    print()
