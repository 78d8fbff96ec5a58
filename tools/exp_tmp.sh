# round 2, session 2, final call: the whole GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s2c9_pytest.txt; cat gpurun_out/s2c9_pytest.txt
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/s2c9_bench.json 2> gpurun_out/s2c9_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c9_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('fwd_plus_bwd_frac'), d['roofline'].get('traffic'), d['intertwiner_loss']['ms_per_iter'], d['gpu_launches_per_step'], {k:(v.get('avg_ms'), v.get('frac')) for k,v in d['kernels'].items()}, {k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items()}, d['e2e']['ms_per_step'], d['ms_each_step'])
PY
