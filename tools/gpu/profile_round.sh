# Round profile set (single GPU): ncu --set full of the two dominant kernels, ncu launch list of a short bench run,
# the full bench line (with cpu_baseline and the strong-scaling block), the reference arm, the secondary configs, and the
# whole GPU test suite.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${tag}_pytest.log 2>&1; tail -n 5 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
ncu --set full --clock-control none --import-source on --target-processes application-only -k regex:"sglShadeKernel|sglVisKernel" -s 8 -c 2 \
    -o gpurun_out/${tag}_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --target-processes application-only -s 80 -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${tag}_ncu_launches.log 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --steps 20 --warmup 3 --no-strong > gpurun_out/${tag}_bench_steps20.json 2> gpurun_out/${tag}_bench_steps20.err
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python tools/bench_configs.py --only c1,c3,c4,c4big,c4full,c5 --out gpurun_out/${tag}_configs.json > gpurun_out/${tag}_configs.log 2>&1
tail -c 900 gpurun_out/${tag}_bench.json; echo; tail -c 300 gpurun_out/${tag}_bench_steps20.json; echo; tail -c 400 gpurun_out/${tag}_bench_reference.json; echo
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_configs.json"))
for k, v in d.items():
    print(k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items()})
PY
