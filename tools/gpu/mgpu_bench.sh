# usage: mgpu_bench.sh N TAG  -- frame-parallel (p2p, nccl) and tile-sharded (p2p) bench lines at N GPUs + the 2-rank parity worker
N=$1; tag=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tests/mgpu_worker.py 2>&1 | grep -E "MGPU_OK|Error|error|assert" | head -5
for mode in "frames p2p" "frames nccl" "tiles p2p"; do
  set -- $mode
  timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 20 --mgpu $1 --gather $2 > gpurun_out/${tag}_n${N}_$1_$2.json 2> gpurun_out/${tag}_n${N}_$1_$2.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${tag}_n${N}_$1_$2.json") if l.startswith("{")][0]
    print("N=$N $1 $2: value %.0f e2e %.0f %s launches %d" % (d["value"], d["e2e"]["value"], d["scaling"], d["gpu_launches"]))
except Exception as e:
    print("N=$N $1 $2 FAILED", e); print(open("gpurun_out/${tag}_n${N}_$1_$2.err").read()[-1500:])
PY
done
