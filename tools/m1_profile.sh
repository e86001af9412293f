#!/bin/bash
# Launch list of ONE single proof (m = 1): ncu per-launch durations of the second of two proofs, grouped by kernel.
# usage: tools/m1_profile.sh [tag]   -> gpurun_out/<tag>_m1_launches.csv + a summary on stdout
tag=${1:-m1}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_m1_launches.csv python tools/prover_perf.py 1 withdraw 2 > /dev/null 2>&1
python - "$tag" <<'PY'
import csv, collections, sys
rows = list(csv.reader(open("gpurun_out/%s_m1_launches.csv" % sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
seq = [(r[ki].split("(")[0], float(r[vi].replace(",", "")), r[gi], r[bi]) for r in rows[start:] if len(r) > vi]
idx = [i for i, s in enumerate(seq) if s[0] == "k_scatter_random"][-1]
last = seq[idx:]
tot, cnt = collections.defaultdict(float), collections.Counter()
for n, v, g, b in last:
    tot[n] += v; cnt[n] += 1
print("launches in one proof:", len(last), "total us: %.1f" % (sum(tot.values()) / 1e3))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("%-28s %4d %9.1f us" % (k, cnt[k], v / 1e3))
print("---- sequence")
for n, v, g, b in last:
    print("%-28s %8.1f us grid %s block %s" % (n, v / 1e3, g, b))
PY
