# round 2, session 2, call 7: verification of the final state: full GPU suite, forward timing, ncu of the default forward, microbench, bench line, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s2c7_pytest.txt; cat gpurun_out/s2c7_pytest.txt
timeout 200 python tools/fwd_ab.py --workload c2 --iters 20 --forms 1,0,3:0:0:1,0 --out gpurun_out/s2c7_fwd_ab_c2.json 2>gpurun_out/s2c7_fwd_ab_c2.err | grep -v '^{"'
timeout 200 python tools/fwd_ab.py --workload c5 --iters 10 --forms 1,0 --out gpurun_out/s2c7_fwd_ab_c5.json 2>gpurun_out/s2c7_fwd_ab_c5.err | grep -v '^{"'
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_fwd_nhwc_sets -c 1 -o gpurun_out/s2c7_ncu_fwd_default -f python tools/fwd_ab.py --iters 1 --forms 0 > gpurun_out/s2c7_ncu.log 2>&1; tail -1 gpurun_out/s2c7_ncu.log
timeout 300 python tools/microbench.py --workload c2 --out gpurun_out/s2c7_microbench_c2.json > gpurun_out/s2c7_microbench.log 2>&1; tail -2 gpurun_out/s2c7_microbench.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/s2c7_bench.json 2> gpurun_out/s2c7_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c7_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('fwd_plus_bwd_frac'), d['intertwiner_loss']['ms_per_iter'], d['gpu_launches_per_step'], {k:(v.get('avg_ms'), v.get('frac')) for k,v in d['kernels'].items()}, {k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items()}, d['e2e']['ms_per_step'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s2c7_launches.csv python bench.py --steps 2 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/s2c7_launches_bench.log 2>&1; tail -c 200 gpurun_out/s2c7_launches_bench.log
