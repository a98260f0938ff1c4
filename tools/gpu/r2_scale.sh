# Round-2 scaling set at N GPUs (gpurun --gpus N): the contract bench line (frame-parallel + strong-scaling block), the
# tile-sharded headline variant, and the secondary configs (configs 3 / 4 tile-sharded, config 5 view-parallel).
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
$TR bench.py --gpus $N --steps 500 --warmup 50 > gpurun_out/r02_bench_n${N}_frames.json 2> gpurun_out/r02_bench_n${N}_frames.err
$TR bench.py --gpus $N --steps 300 --warmup 30 --mgpu tiles --no-strong > gpurun_out/r02_bench_n${N}_tiles.json 2> gpurun_out/r02_bench_n${N}_tiles.err
$TR tools/bench_configs.py --gather ${3:-p2p} --only ${2:-c3,c4big,c5} --out gpurun_out/r02_configs_n${N}.json > gpurun_out/r02_configs_n${N}.log 2>&1
python - <<PY
import json
for f in ("frames", "tiles"):
    try:
        d = json.load(open("gpurun_out/r02_bench_n${N}_%s.json" % f))
        print(f, "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), {k: round(v["units_per_s"], 1) for k, v in (d.get("strong_scaling") or {}).items() if isinstance(v, dict) and "units_per_s" in v})
    except Exception as e:
        print(f, "failed", e)
try:
    d = json.load(open("gpurun_out/r02_configs_n${N}.json"))
    print({k: round(v["units_per_s"], 1) for k, v in d.items()})
except Exception as e:
    print("configs failed", e)
PY
for f in gpurun_out/r02_bench_n${N}_frames.err gpurun_out/r02_configs_n${N}.log; do tail -n 2 "$f"; done; true
