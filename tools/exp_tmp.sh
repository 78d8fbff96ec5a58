# round 2, session 2, call 2: forward schedule sweep (shape x chunk x pairing), parity of the new schedules, ncu of the default
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_roi_align_gpu.py -m gpu -x -q -k "forward or crop_sets or full_size or golden or known" 2>&1 | tail -4 > gpurun_out/s2c2_pytest.txt; cat gpurun_out/s2c2_pytest.txt
timeout 400 python tools/fwd_ab.py --workload c2 --iters 20 --forms 1,3:1:1,3:1:2,3:2:1,3:2:2,3:3:2,3:4:2,3:5:2,3:6:2,4:2:2,4:3:2,4:4:2,5:2:2,5:4:2,6:2:2,6:3:2,2:2:2,3:2:2,3:3:2 --out gpurun_out/s2c2_fwd_ab_c2.json 2>gpurun_out/s2c2_fwd_ab_c2.err | grep -v '^{"' 
timeout 300 python tools/fwd_ab.py --workload c5 --iters 10 --forms 1,3:1:1,3:2:2,3:3:2,3:4:2,4:3:2,5:2:2,6:2:2 --out gpurun_out/s2c2_fwd_ab_c5.json 2>gpurun_out/s2c2_fwd_ab_c5.err | grep -v '^{"'
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_fwd_nhwc_sets -c 1 -o gpurun_out/s2c2_ncu_fwd_lean -f python tools/fwd_ab.py --iters 1 --forms 0 > gpurun_out/s2c2_ncu.log 2>&1; tail -2 gpurun_out/s2c2_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crop_fwd_nhwc_sets -c 1 -o gpurun_out/s2c2_ncu_fwd_lean_nopair -f python tools/fwd_ab.py --iters 1 --forms 3:1:1 > gpurun_out/s2c2_ncu2.log 2>&1; tail -2 gpurun_out/s2c2_ncu2.log
