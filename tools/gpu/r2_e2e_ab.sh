# e2e diagnostics on config 2: environment switches given as arguments (each run twice, interleaved), with the e2e region
# seen from the device and from the host; then the streams parity tests
mkdir -p gpurun_out
for rep in 1 2; do
for sw in "X=0" "$@"; do
env $sw python bench.py --steps 500 --warmup 50 --no-cpu-baseline --no-strong 2>gpurun_out/quick_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_frame']
print('$sw c2 value %.0f e2e %.0f | e2e ms/step: wall %.4f device %.4f drained %.4f host-blocked %.4f | host_submit %.3f numa %s' % (d['value'], d['e2e']['value'], 1e3/d['e2e']['value'], d['e2e_device_ms_per_step'], d['e2e_wall_ms_until_main_stream_drained_per_step'], d['e2e_host_blocked_ms_per_step'], d['host_submit_ms_per_step'], d['host_numa']))"
done
done
python -m pytest tests/test_streams_gpu.py -m gpu -q -x 2>&1 | tail -3
# PCIe read-back rate with the GPU busy: 8.3 MB pinned copies on a side stream while a long kernel sequence runs
python - <<PY
import torch, time
n = 8294400
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
a = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
s2 = torch.cuda.Stream()
for busy in (0, 1):
    torch.cuda.synchronize()
    if busy:
        for _ in range(40): b.copy_(a)          # ~2 GB of HBM traffic each: the SMs/HBM stay busy for tens of ms
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s2):
        e0.record()
        for _ in range(20): h.copy_(d, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    print("D2H 8.3 MB x20, GPU %s: %.1f GB/s" % ("busy" if busy else "idle", n * 20 / e0.elapsed_time(e1) / 1e6))
PY
