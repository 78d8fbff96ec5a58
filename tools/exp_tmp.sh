# round 2, session 2, call 3: fused loss head + final forward default: full GPU suite, forward timing, ncu of the default forward, bench line, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2c3_pytest.txt; cat gpurun_out/s2c3_pytest.txt
timeout 300 python tools/fwd_ab.py --workload c2 --iters 20 --forms 1,0,1,0 --out gpurun_out/s2c3_fwd_ab_c2.json 2>gpurun_out/s2c3_fwd_ab_c2.err | grep -v '^{"'
timeout 300 python tools/fwd_ab.py --workload c5 --iters 10 --forms 1,0 --out gpurun_out/s2c3_fwd_ab_c5.json 2>gpurun_out/s2c3_fwd_ab_c5.err | grep -v '^{"'
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_fwd_nhwc_sets -c 1 -o gpurun_out/s2c3_ncu_fwd_default -f python tools/fwd_ab.py --iters 1 --forms 0 > gpurun_out/s2c3_ncu.log 2>&1; tail -2 gpurun_out/s2c3_ncu.log
timeout 600 python bench.py > gpurun_out/s2c3_bench.json 2> gpurun_out/s2c3_bench.err; tail -c 600 gpurun_out/s2c3_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c3_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['intertwiner_loss'], d['gpu_launches_per_step'], {k:v.get('avg_ms') for k,v in d['kernels'].items()}, {k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s2c3_launches.csv python bench.py --steps 2 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/s2c3_launches_bench.log 2>&1; tail -c 300 gpurun_out/s2c3_launches_bench.log
