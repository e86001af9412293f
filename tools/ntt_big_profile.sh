#!/bin/bash
# ncu --set full of the two tile passes of one 2^22-point NTT (the sweep size the TMA clause of the north star is about)
# usage: tools/ntt_big_profile.sh <tag>
tag=$1
cat > /tmp/ntt_big.py <<'PY'
import os, sys
import numpy as np, torch
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, ROOT + "/zkos-monorepo_b200", ROOT + "/tests"]
import zkgpu
from zkgpu.gpu_backend import GpuBackend as F
zkgpu.init(0)
log_n, batch = 22, 4
n = 1 << log_n
x = torch.from_numpy(F.random(5, 1 << 14).view(np.int64)).cuda().repeat((n * batch) >> 14, 1).contiguous()
scratch = torch.empty_like(x)
from pyref import omega_for, int_to_limbs
w = F.to_mont(int_to_limbs([omega_for(log_n)]))[0]
for _ in range(3):
    zkgpu.ntt_batch_dev(x.data_ptr(), w, log_n, batch, scratch.data_ptr())
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:k_ntt_tile -s 2 -c 2 -o gpurun_out/${tag} python /tmp/ntt_big.py > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python - "$tag" <<'PY'
import csv, sys
rows = list(csv.reader(open("gpurun_out/%s_raw.csv" % sys.argv[1])))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: r[i][:48] for i in idx})
PY
