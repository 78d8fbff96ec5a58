# round 2, session 2, call 1: full GPU suite, forward formulation sweep, ncu of the lean forward, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s2c1_pytest.txt; cat gpurun_out/s2c1_pytest.txt
timeout 300 python tools/fwd_ab.py --workload c2 --iters 20 --forms 1,0,3,4,5,6,1,0 --out gpurun_out/s2c1_fwd_ab_c2.json 2>gpurun_out/s2c1_fwd_ab_c2.err | tail -12
timeout 300 python tools/fwd_ab.py --workload c5 --iters 10 --forms 1,0,3,4,5,6 --out gpurun_out/s2c1_fwd_ab_c5.json 2>gpurun_out/s2c1_fwd_ab_c5.err | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_fwd_nhwc_sets -c 2 -o gpurun_out/s2c1_ncu_fwd_lean -f python tools/fwd_ab.py --iters 1 --forms 0 > gpurun_out/s2c1_ncu.log 2>&1; tail -2 gpurun_out/s2c1_ncu.log
timeout 600 python bench.py > gpurun_out/s2c1_bench.json 2> gpurun_out/s2c1_bench.err; tail -c 1500 gpurun_out/s2c1_bench.json
