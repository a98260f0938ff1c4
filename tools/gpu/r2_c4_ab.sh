for sw in X=0 SGL_NO_EARLY_VIS=1 SGL_RING=3 "SGL_NO_EARLY_VIS=1 SGL_RING=3"; do
env $sw python tools/bench_configs.py --only c4,c4big --out gpurun_out/ab_configs.json > gpurun_out/ab_configs.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/ab_configs.json"))
for k, v in d.items():
    print("$sw", k, round(v["units_per_s"], 1), "ms", round(v["ms_per_step"], 3), "sum of kernels", round(sum(v["kernel_ms_per_step"].values()), 3))
PY
done
