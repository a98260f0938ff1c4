# N = 2 validation of the final build (gpurun --gpus 2): the real 2-rank tests, then the contract bench line frame-parallel
# (direct peer stores and NCCL) and tile-sharded, condensed
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 500 --warmup 50 > gpurun_out/r02_bench_n2_frames.json 2> gpurun_out/r02_bench_n2_frames.err; tail -c 400 gpurun_out/r02_bench_n2_frames.err
timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 30 --gather nccl --no-strong > gpurun_out/r02_bench_n2_frames_nccl.json 2> gpurun_out/r02_bench_n2_frames_nccl.err; tail -c 400 gpurun_out/r02_bench_n2_frames_nccl.err
timeout 300 $TR bench.py --gpus 2 --steps 300 --warmup 30 --mgpu tiles --no-strong > gpurun_out/r02_bench_n2_tiles.json 2> gpurun_out/r02_bench_n2_tiles.err; tail -c 400 gpurun_out/r02_bench_n2_tiles.err
for f in gpurun_out/r02_bench_n2_*.json; do python - <<PY
import json
for l in open("$f"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$f".split("/")[-1], "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), d["scaling"], "early_vis/step", d.get("early_vis_per_step"), d.get("early_vis_per_step_e2e"),
              {k: round(v["units_per_s"], 1) for k, v in (d.get("strong_scaling") or {}).items() if isinstance(v, dict) and "units_per_s" in v}, (d.get("strong_scaling") or {}).get("error"))
PY
done
