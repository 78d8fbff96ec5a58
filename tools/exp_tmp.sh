timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_b.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","host_enqueue_ms_per_step"): print(k, d[k])
print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("note","traffic_source")})
for k,v in d["kernels"].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
