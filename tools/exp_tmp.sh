# round 2, session 2, call 6: ticket schedule of the lean forward: parity, sweep over units per draw / shapes, ncu, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_roi_align_gpu.py -m gpu -x -q -k "forward or crop_sets or full_size or golden" 2>&1 | tail -3
timeout 400 python tools/fwd_ab.py --workload c2 --iters 20 --forms 1,3:0:0:1,3:1:0:2,3:2:0:2,3:3:0:2,3:4:0:2,3:5:0:2,3:6:0:2,4:2:0:2,4:5:0:2,6:2:0:2,6:5:0:2,2:2:0:2,5:2:0:2,3:2:1:2,3:0:0:1,3:2:0:2 --out gpurun_out/s2c6_fwd_ab_c2.json 2>gpurun_out/s2c6_fwd_ab_c2.err | grep -v '^{"'
timeout 300 python tools/fwd_ab.py --workload c5 --iters 10 --forms 1,3:0:0:1,3:1:0:2,3:2:0:2,3:5:0:2,4:2:0:2,6:2:0:2 --out gpurun_out/s2c6_fwd_ab_c5.json 2>gpurun_out/s2c6_fwd_ab_c5.err | grep -v '^{"'
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_fwd_nhwc_sets -c 1 -o gpurun_out/s2c6_ncu_fwd_tickets -f python tools/fwd_ab.py --iters 1 --forms 0 > gpurun_out/s2c6_ncu.log 2>&1; tail -1 gpurun_out/s2c6_ncu.log
timeout 600 python bench.py --no-other-workloads > gpurun_out/s2c6_bench.json 2> gpurun_out/s2c6_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c6_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['intertwiner_loss']['ms_per_iter'], d['intertwiner_loss']['graphed'], d['intertwiner_loss']['graph_error'], d['gpu_launches_per_step'], {k:v.get('avg_ms') for k,v in d['kernels'].items()}, d['ms_each_step'])
PY
