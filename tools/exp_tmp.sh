timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_roi_align_gpu.py tests/test_ops_gpu.py -m gpu -x -q -k "launcher or backward_matches_oracle or roi_pool or proposal" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-workloads > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_r02_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02_n2.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","host_enqueue_ms_per_step","n_gpus"): print(k, d[k])
print(d["config"]["step"][:200]); print("e2e", d["e2e"]); print("nvlink", d.get("nvlink")); print("loss", d["intertwiner_loss"]["ms_per_iter"])
PY
