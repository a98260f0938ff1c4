# deferred (visibility + shading) path against the fused tile kernel (coverage -> depth -> shading -> resolve with the tile's
# depth / owners / colour in registers, nothing but the attachments in memory) on the large-frame configs
for sw in X=0 SGL_FORCE_FUSED=1; do
env $sw python tools/bench_configs.py --only ${1:-c3,c4} --out gpurun_out/fused_ab.json > gpurun_out/fused_ab.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/fused_ab.json"))
for k, v in d.items():
    print("$sw", k, round(v["units_per_s"], 1), "ms", round(v["ms_per_step"], 3), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items() if any(x in n for x in ("Vis", "Shade", "Raster"))})
PY
done
