timeout 600 python bench.py --steps 10 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_c.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","host_enqueue_ms_per_step","gpu_launches_per_step"): print(k, d[k])
print("loss", d["intertwiner_loss"]["ms_per_iter"]); print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["host_buffers"])
for k,v in d["kernels"].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_v1.csv python bench.py --steps 2 --warmup 3 --mode eager --no-other-workloads --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02_launches_bench_v1.csv")) if len(r)>10 and r[0].isdigit()]
print(len(rows))
agg=collections.OrderedDict()
for r in rows[-260:]:
    n=r[4][:70]; agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=float(r[-1])/1e3
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:30]: print("%4d %9.1f us  %s"%(c,t,n))
PY
