# usage: mgpu_quick.sh N TAG -- default bench line (frame-parallel, direct peer stores) and the tile-sharded one at N GPUs
N=$1; tag=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
for mode in "frames p2p" "tiles p2p"; do
  set -- $mode
  timeout 400 $TR bench.py --gpus $N --steps 200 --warmup 20 --mgpu $1 --gather $2 > gpurun_out/${tag}_n${N}_$1_$2.json 2> gpurun_out/${tag}_n${N}_$1_$2.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${tag}_n${N}_$1_$2.json") if l.startswith("{")][0]
    print("N=$N $1 $2: value %.0f e2e %.0f %s launches %d gather=%s" % (d["value"], d["e2e"]["value"], d["scaling"], d["gpu_launches"], d["config"]["gather"][:40]))
except Exception as e:
    print("N=$N $1 $2 FAILED", e); print(open("gpurun_out/${tag}_n${N}_$1_$2.err").read()[-2000:])
PY
done
