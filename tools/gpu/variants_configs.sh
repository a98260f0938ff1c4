# secondary configs for one compile-time variant against the shipped build: variants_configs.sh NAME [cases]
for v in "" "$1"; do
  if [ -n "$v" ]; then export SGL_LIB_DIR=$PWD/softglrender_b200/lib_variants/$v; else unset SGL_LIB_DIR; fi
  python tools/bench_configs.py --only ${2:-c1,c3,c4,c4big,c5} --out gpurun_out/var_configs.json > gpurun_out/var_configs.log 2>&1
  python - <<PY
import json
d = json.load(open("gpurun_out/var_configs.json"))
for k, v in d.items():
    print("${v:-base}", k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items() if "Vis" in n or "Shade" in n})
PY
done
