timeout 600 python -m pytest tests/test_roi_align_gpu.py -m gpu -x -q -k "bulk_copy_staged" 2>&1 | tail -4
for f in 3 0; do FI_FWD_FORM=$f timeout 200 python bench.py --steps 5 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_fwd$f.json 2>gpurun_out/bench_fwd$f.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_fwd$f.json").read().strip().splitlines()[-1])
print("fwd_form $f", "ms/step", round(d["ms_per_step"],4), {k:round(v["avg_ms"],4) for k,v in d["kernels"].items()})
PY
done
