#!/bin/bash
# ncu --set full of selected kernels of a single proof (m = 1): tools/ncu_m1.sh <tag> <kernel regex> [count]
tag=$1; k=$2; cnt=${3:-4}
ncu --set full --clock-control none --import-source on -k "regex:$k" -s 6 -c $cnt -o gpurun_out/${tag} python tools/prover_perf.py 1 withdraw 2 > /dev/null 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python - "$tag" <<'PY'
import csv, sys
rows = list(csv.reader(open("gpurun_out/%s_raw.csv" % sys.argv[1])))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "launch__occupancy_limit_registers", "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: r[i][:60] for i in idx})
PY
