#!/bin/bash
# A/B runs of the mixed stream at the per-rank load of an N-GPU job, on ONE GPU: tools/ab_mixed.sh <tag> <requests> "ENV=1" ...
tag=$1; req=$2; shift 2
i=0
for cfg in "default" "$@"; do
  envs=""; [ "$cfg" != "default" ] && envs="$cfg"
  env $envs python bench.py --workload mixed --requests $req --no-cpu-baseline > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$cfg" gpurun_out/${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("%-44s value %.1f ms_per_step %.1f" % (sys.argv[1], d["value"], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  i=$((i+1))
done
