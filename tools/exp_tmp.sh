timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r02_a.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02_a.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","host_enqueue_ms_per_step","gpu_launches_per_step","ms_each_step"): print(k, d[k])
    print("step:", d["config"]["step"][:120])
    print("e2e", d["e2e"]); print("loss", d["intertwiner_loss"]); print("eager", d["eager"])
    print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("note","traffic_source")})
    for k,v in d["kernels"].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
    print("others", d["other_workloads"]); print("cpu", d.get("cpu_baseline"))
except Exception as e: print("ERR", e)
PY
timeout 900 python -m pytest tests/test_roi_align_gpu.py -m gpu -x -q -k "full_size_vs_compiled" 2>&1 | tail -8
