"""Key metrics per kernel launch from an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`).
usage: python tests/runs/ncu_summary.py report.ncu-rep "header" > profiles/xxx_ncu_full.txt"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
idx = {n: i for i, n in enumerate(h)}
print(header)
names = []
for r in data:
    n = re.sub(r"\(.*$", "", r[idx["Kernel Name"]]).replace("void ", "").replace("samq::<unnamed>::", "").replace("samq::", "")
    names.append(n)
print(f"{'Kernel Name':72s} [] " + " | ".join(names))
for m in METRICS:
    if m not in idx:
        continue
    print(f"{m:72s} [{units[idx[m]]}] " + " | ".join(r[idx[m]] for r in data))
