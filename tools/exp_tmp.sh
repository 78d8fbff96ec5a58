timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_roi_align_gpu.py -m gpu -x -q -k "roi_pool or formulations or in_place or device_counts or golden or overflow" 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_enumerate|tile_prep" -c 6 --csv --log-file gpurun_out/r02_launches_enum3.csv python tools/bwd_ab.py --iters 2 > /dev/null 2>&1; grep -o '"[a-z_:A-Z ]*\(bin_enumerate\|tile_prep\)[^"]*".*' gpurun_out/r02_launches_enum3.csv | awk -F'","' '{print $1, $NF}' | tail -4
timeout 120 python tools/bwd_ab.py --iters 12 --plan-in-backward --out gpurun_out/ab_enum3_all.json > /dev/null 2>gpurun_out/ab_enum3.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_enum3_all.json")); print("all", round(d["bwd_ms_median"],4), round(d["bwd_ms_min"],4), d["fingerprints"][0][0])
PY
