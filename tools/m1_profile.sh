ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_m1.csv python tools/prover_perf.py 1 withdraw 2 > /dev/null 2>&1; python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches_m1.csv")))
for i,r in enumerate(rows):
    if r and r[0]=="ID": hdr=r; start=i+1; break
ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
seq=[(r[ki].split("(")[0], float(r[vi].replace(",","")), r[gi]) for r in rows[start:] if len(r)>vi]
idx=[i for i,s in enumerate(seq) if s[0]=="k_scatter_random"][-1]
last=seq[idx:]
tot=collections.defaultdict(float); cnt=collections.Counter()
for n,v,g in last: tot[n]+=v; cnt[n]+=1
print("launches in one proof:", len(last), "total us:", sum(tot.values())/1e3)
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:12]: print("%-22s %4d %9.1f us" % (k,cnt[k],v/1e3))
print([ (n,round(v/1e3),g) for n,v,g in last if n in ("k_msm_buckets_split","k_msm_reduce","k_normalize","k_msm_heavy")])
PY
