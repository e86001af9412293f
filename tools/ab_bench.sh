#!/bin/bash
# A/B runs of bench.py under different run-time switches: tools/ab_bench.sh <tag> "ENV=1 ENV2=2" "ENV=3" ...
# Each configuration's JSON line goes to gpurun_out/<tag>_<i>.json; a one-line summary per configuration is printed.
tag=$1; shift
i=0
for cfg in "default" "$@"; do
  envs=""; [ "$cfg" != "default" ] && envs="$cfg"
  env $envs python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$cfg" gpurun_out/${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-40s value %.1f e2e %.1f p50 %.2f ms  kernels %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["single_proof_p50_ms"],
          {k: round(v, 1) for k, v in d["kernel_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  i=$((i+1))
done
