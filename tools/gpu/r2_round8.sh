# visibility kernel with one copy of the per-primitive body: parity suites, config 2, secondary configs
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py tests/test_streams_gpu.py tests/test_overflow_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -3
bash tools/gpu/r2_c2_ab.sh 2>&1 | head -2
python tools/bench_configs.py --only c1,c3,c4,c4big,c5 --out gpurun_out/ab_configs.json > gpurun_out/ab_configs.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/ab_configs.json"))
for k, v in d.items():
    print(k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items() if "Vis" in n or "Shade" in n or "Raster" in n})
PY
