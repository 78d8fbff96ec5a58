# round 2, session 2, call 5 (2 GPUs): NCCL parity tests of the sharded step with the kernel merge, N=2 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-workloads > gpurun_out/s2c5_bench_n2.json 2> gpurun_out/s2c5_bench_n2.err; python - <<'PY'
import json
for l in open('gpurun_out/s2c5_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value'], d['config'].get('step'), d['intertwiner_loss'], d['gpu_launches_per_step'], d.get('nvlink',{}).get('class_stats_allreduce_peer_kernel'))
PY
tail -c 400 gpurun_out/s2c5_bench_n2.err
