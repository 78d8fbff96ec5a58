# round 2, session 2, call 4: loss-head graph diagnosis, fused head v2 (fused ReLU epilogue, column sums), bench
mkdir -p gpurun_out
timeout 300 python tools/diag_loss_graph.py > gpurun_out/s2c4_diag.txt 2>&1; grep -v Warning gpurun_out/s2c4_diag.txt | tail -30
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/s2c4_bench.json 2> gpurun_out/s2c4_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c4_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['intertwiner_loss'], d['gpu_launches_per_step'], {k:v.get('avg_ms') for k,v in d['kernels'].items()}, {k:v.get('ms_per_step') for k,v in d.get('other_workloads',{}).items()})
PY
