# single-sample heavy tiles as four quarter-tile CTAs (four triangles per pixel at a time): parity, then A/B on the secondary configs
python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py tests/test_streams_gpu.py tests/test_overflow_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -3
for mode in 0 1; do
  SGL_NO_SPLIT1=$mode python tools/bench_configs.py --only ${1:-c1,c3,c4,c4big,c4full,c5} --out gpurun_out/r02_split1_off$mode.json > gpurun_out/r02_split1_$mode.log 2>&1
  tail -n 1 gpurun_out/r02_split1_$mode.log | cut -c1-200
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_split1_off$mode.json"))
for k, v in d.items():
    print("nosplit1=$mode", k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items()})
PY
done
