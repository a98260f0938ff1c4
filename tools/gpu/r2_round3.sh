# split rule relative to owned tiles + line-visit reorder: parity suites, config 2 condensed, one-rank-of-eight simulation
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -3
bash tools/gpu/r2_c2_ab.sh 2>&1 | head -1
SKIP_TESTS=1 SGL_NO_LAZY_VARYINGS=0 python tools/bench_configs.py --only c3,c4big --as-rank 0/8 --out gpurun_out/r02_shard_sim_r0.json > gpurun_out/r02_shard_sim_r0.log 2>&1
SKIP_TESTS=1 python tools/bench_configs.py --only c4big --as-rank 3/8 --out gpurun_out/r02_shard_sim_r3.json > gpurun_out/r02_shard_sim_r3.log 2>&1
python - <<PY
import json
for f in ("r0", "r3"):
    d = json.load(open("gpurun_out/r02_shard_sim_%s.json" % f))
    for k, v in d.items():
        print(f, k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items()})
PY
