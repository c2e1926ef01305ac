"""Turns `ncu -i X.ncu-rep --page raw --csv` into the short summary kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "label" >> profiles/rNN_ncu_summary.txt"""
import csv
import subprocess
import sys

rep, label = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"]
for r in rows[2:]:
    print(f"== {label}: {r[idx['Kernel Name']][:70]}")
    for k in keys:
        if k in idx:
            print(f"   {k:82s} {r[idx[k]]:>16s} {units[idx[k]]}")
