timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -6
for w in c2 c3; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w", "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), {k:round(v["avg_ms"],4) for k,v in d["kernels"].items()}, "loss", round(d["intertwiner_loss"]["ms_per_iter"],4))
PY
done
