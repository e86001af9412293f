#!/bin/bash
# One `ncu --set full` capture per dominant kernel of the prover (run under gpurun; CSV export on the box).
set -u
TAG=${1:-r01}
i=0
for spec in 'k_msm_buckets$:2:8' 'k_eval_h:0:1' 'k_msm_reduce$:2:1' 'k_msm_sort_smem:2:1' 'k_ntt_cluster8:1:2' 'k_coset_combine:0:1' 'k_msm_heavy:2:1' 'k_batch_inverse:0:1'; do
    IFS=: read -r name skip count <<< "$spec"
    ncu --set full --clock-control none --import-source on -k "regex:$name" -s "$skip" -c "$count" -f -o /tmp/prof_${TAG}_$i \
        python tools/prover_perf.py 128 withdraw 1 > gpurun_out/ncu_k_${TAG}_$i.log 2>&1
    ncu -i /tmp/prof_${TAG}_$i.ncu-rep --page raw --csv > gpurun_out/ncu_k_${TAG}_$i.csv 2>/dev/null
    i=$((i+1))
done
python - <<PY
import csv, glob
first = True
with open("gpurun_out/ncu_kernels_${TAG}_raw.csv", "w", newline="") as out:
    w = csv.writer(out)
    for f in sorted(glob.glob("gpurun_out/ncu_k_${TAG}_*.csv")):
        rows = list(csv.reader(open(f)))
        if len(rows) < 3: continue
        if first: w.writerow(rows[0]); w.writerow(rows[1]); first = False
        for r in rows[2:]: w.writerow(r)
PY
rm -f gpurun_out/ncu_k_${TAG}_*.csv
