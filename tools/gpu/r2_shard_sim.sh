# What one rank of an 8-way tile-sharded run does, on ONE GPU (no exchange): lazy varyings + owned-only emission vs up-front varyings
[ -n "$SKIP_TESTS" ] || python -m pytest tests/test_multigpu_gpu.py tests/test_parity_gpu.py tests/test_parity_configs_gpu.py -m gpu -q -x 2>&1 | tail -3
for mode in 0 1; do
  SGL_NO_LAZY_VARYINGS=$mode python tools/bench_configs.py --only ${1:-c3,c4big,c4full} --as-rank 0/8 --out gpurun_out/r02_shard_sim_nolazy$mode.json > gpurun_out/r02_shard_sim_$mode.log 2>&1
  tail -n 2 gpurun_out/r02_shard_sim_$mode.log | cut -c1-300
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_shard_sim_nolazy$mode.json"))
for k, v in d.items():
    print("nolazy=$mode", k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items()})
PY
done
