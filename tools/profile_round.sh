#!/bin/bash
# Round profile bundle (run under gpurun): plain bench, ncu launch list of the same command, full-metric
# captures of the dominant kernels (exported to CSV on the box: .ncu-rep files are too large to bring back).
set -u
TAG=${1:-r01_final}
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
KREG='regex:k_(msm|lat|ntt|eval|coset|pull|perm|batch|poly|lincomb|kate|chacha|scatter_random|normalize|lookup|sub_low)'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 400 -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_${TAG}.log 2>&1
python tools/prover_perf.py 128 withdraw 1 > gpurun_out/pp_${TAG}.log 2>&1 || exit 1
ncu --set full --clock-control none -k 'regex:k_msm_buckets$|k_ntt_tile|k_ntt_cluster8|k_eval_h|k_msm_reduce$|k_msm_sort_smem|k_coset_combine' -s 30 -c 16 -o /tmp/prof_${TAG} \
    python tools/prover_perf.py 128 withdraw 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | head -20
