# round 2, session 2: N=2 line with the final kernels and the device-side rendezvous
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/s2c10_bench_n2.json 2> gpurun_out/s2c10_bench_n2.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c10_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value'], d['ms_per_step_by_rank'], d['ms_each_step'], d['intertwiner_loss']['ms_per_iter'], d['gpu_launches_per_step'])
PY
tail -c 300 gpurun_out/s2c10_bench_n2.err
