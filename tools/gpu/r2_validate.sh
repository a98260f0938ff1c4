# Validation of the current build on one GPU: full GPU suite, smoke(), condensed bench line, then what one rank of an
# 8-way tile-sharded run does (lazy varyings on / off) on configs 3 and 4 (2 M triangles).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/val_pytest.log 2>&1; tail -n 6 gpurun_out/val_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-strong 2>gpurun_out/quick_bench.err | tee gpurun_out/quick_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_frame']
print('c2 value %.0f e2e %.0f host %.3f | ' % (d['value'], d['e2e']['value'], d['host_submit_ms_per_step']) + ' '.join('%s=%.0f' % (n.replace('sgl','').replace('Kernel',''), t*1e3) for n,t in sorted(k.items())))"
tail -n 3 gpurun_out/quick_bench.err
timeout 400 python tools/bench_configs.py --only c1,c3,c4,c4big,c5 --out gpurun_out/val_configs.json > gpurun_out/val_configs.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/val_configs.json"))
for k, v in d.items():
    print(k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items()})
PY
SKIP_TESTS=1 timeout 500 bash tools/gpu/r2_shard_sim.sh c3,c4big
