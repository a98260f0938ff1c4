# early visibility + depth-pass renaming: streams suite, parity suites, configs with the switches
mkdir -p gpurun_out
python -m pytest tests/test_streams_gpu.py -m gpu -q -x 2>&1 | tail -5
python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py tests/test_overflow_gpu.py tests/test_multigpu_gpu.py tests/test_viewer_integration_gpu.py -m gpu -q -x 2>&1 | tail -3
bash tools/gpu/r2_c2_ab.sh SGL_NO_RENAME=1 2>&1 | head -2
for sw in X=0 SGL_NO_RENAME=1 SGL_NO_EARLY_VIS=1; do
env $sw python tools/bench_configs.py --only c1,c5 --out gpurun_out/ab_configs.json > gpurun_out/ab_configs.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/ab_configs.json"))
for k, v in d.items():
    print("$sw", k, round(v["units_per_s"], 1))
PY
done
