# A/B of an environment switch on the secondary configs: r2_ab.sh "<ENV=1>" "<cases>"   (first line = switch off)
python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py tests/test_streams_gpu.py tests/test_overflow_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -3
for sw in "X=0" $1; do
  env $sw python tools/bench_configs.py --only ${2:-c1,c3,c4,c4big,c5} --out gpurun_out/r02_ab.json > gpurun_out/r02_ab.log 2>&1
  tail -n 1 gpurun_out/r02_ab.log | cut -c1-160
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_ab.json"))
for k, v in d.items():
    print("$sw", k, round(v["units_per_s"], 1), {n.replace("sgl", "").replace("Kernel", ""): round(t * 1e3) for n, t in v["kernel_ms_per_step"].items()})
PY
done
python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-strong 2>gpurun_out/quick_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_frame']
print('c2 value %.0f e2e %.0f host %.3f | ' % (d['value'], d['e2e']['value'], d['host_submit_ms_per_step']) + ' '.join('%s=%.0f' % (n.replace('sgl','').replace('Kernel',''), t*1e3) for n,t in sorted(k.items())))"
