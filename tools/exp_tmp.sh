# round 2, session 2, call 8: small-M GEMM, collapse with more loads in flight, column sums: parity + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_roi_align_gpu.py tests/test_reference_pins.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-other-workloads > gpurun_out/s2c8_bench.json 2> gpurun_out/s2c8_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c8_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('fwd_plus_bwd_frac'), d['intertwiner_loss']['ms_per_iter'], d['intertwiner_loss']['gpu_vs_cpu_port_abs_diff'], d['gpu_launches_per_step'], {k:(v.get('avg_ms'), v.get('frac')) for k,v in d['kernels'].items()}, d['ms_each_step'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s2c8_launches.csv python bench.py --steps 2 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/s2c8_launches_bench.log 2>&1; tail -c 100 gpurun_out/s2c8_launches_bench.log
