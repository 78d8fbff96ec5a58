timeout 900 python -m pytest tests/test_roi_align_gpu.py -m gpu -x -q 2>&1 | tail -15
for c in 0 2; do
  FI_PIX_CFG=$c timeout 120 python tools/bwd_ab.py --iters 12 --out gpurun_out/exp2_c${c}_run.json > /dev/null 2>>gpurun_out/exp2.err || echo "FAIL c$c"
  FI_PIX_CFG=$c timeout 120 python tools/bwd_ab.py --iters 12 --plan-in-backward --out gpurun_out/exp2_c${c}_all.json > /dev/null 2>>gpurun_out/exp2.err || echo "FAIL c$c"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/exp2_c*.json")):
    d=json.load(open(f)); print(f, "median %.4f min %.4f"%(d["bwd_ms_median"], d["bwd_ms_min"]), d["fingerprints"][0][0])
PY
FI_PIX_CFG=2 ncu --set full --clock-control none --import-source on -k regex:"pix_accumulate" -s 2 -c 1 -o gpurun_out/r02_pix_v3 python tools/bwd_ab.py --iters 1 > gpurun_out/ncu_pix_v3.log 2>&1
tail -2 gpurun_out/exp2.err
