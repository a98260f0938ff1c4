# usage: mgpu_gather_modes.sh N TAG [worker] -- frame-parallel bench at N GPUs with the three gather forms
N=$1; tag=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
if [ -n "$3" ]; then timeout 300 $TR tests/mgpu_worker.py 2>&1 | grep -E "MGPU_OK|Error|error|assert" | head -5 | cut -c1-600; fi
for g in p2p dma nccl; do
  timeout 400 $TR bench.py --gpus $N --steps 200 --warmup 20 --gather $g > gpurun_out/${tag}_n${N}_frames_$g.json 2> gpurun_out/${tag}_n${N}_frames_$g.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${tag}_n${N}_frames_$g.json") if l.startswith("{")][0]
    print("N=$N frames $g: value %.0f e2e %.0f launches %d d2h %d gather=%s" % (d["value"], d["e2e"]["value"], d["gpu_launches"], d["e2e"]["d2h_bytes_per_step"], d["config"]["gather"][:30]))
except Exception as e:
    print("N=$N $g FAILED", e); print(open("gpurun_out/${tag}_n${N}_frames_$g.err").read()[-2000:])
PY
done
