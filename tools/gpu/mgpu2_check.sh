mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r1c_n1.json 2> gpurun_out/r1c_n1.err; tail -c 600 gpurun_out/r1c_n1.err
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r1c_n2_frames_p2p.json 2> gpurun_out/r1c_n2_frames_p2p.err; tail -c 600 gpurun_out/r1c_n2_frames_p2p.err
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 20 --gather nccl > gpurun_out/r1c_n2_frames_nccl.json 2> gpurun_out/r1c_n2_frames_nccl.err; tail -c 600 gpurun_out/r1c_n2_frames_nccl.err
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 20 --mgpu tiles > gpurun_out/r1c_n2_tiles_p2p.json 2> gpurun_out/r1c_n2_tiles_p2p.err; tail -c 600 gpurun_out/r1c_n2_tiles_p2p.err
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 20 --mgpu tiles --gather nccl > gpurun_out/r1c_n2_tiles_nccl.json 2> gpurun_out/r1c_n2_tiles_nccl.err; tail -c 600 gpurun_out/r1c_n2_tiles_nccl.err
for f in gpurun_out/r1c_*.json; do echo $f; python -c "
import json,sys
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e'], d['scaling'], d['config'].get('gather'), d['gpu_launches'])
"; done
