# The GPU calls behind the final state of round 2 (run under `gpurun -- 'bash tools/exp_tmp.sh'`; outputs in gpurun_out/).
mkdir -p gpurun_out
# 1. the whole GPU suite, smoke(), the default bench line
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
# 2. forward formulations / schedules, each bit-compared with the round-1 unit (profiles/r02_fwd_ab_*.json)
timeout 300 python tools/fwd_ab.py --workload c2 --iters 20 --forms 1,0,3:0:0:1,3:2:0:2,3:5:0:2 --out gpurun_out/fwd_ab_c2.json | grep -v '^{"'
# 3. ncu of the forward as shipped (profiles/r02_ncu_fwd_lean_v4_tickets_default.txt via tools/ncu_summary.py) and the launch list of a step
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_fwd_nhwc_sets -c 1 -o gpurun_out/ncu_fwd -f python tools/fwd_ab.py --iters 1 --forms 0 > gpurun_out/ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-other-workloads --no-cpu-baseline > /dev/null 2>&1
# 4. two ranks (gpurun --gpus 2): NCCL parity tests and the N = 2 line
# timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q
# timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-workloads
