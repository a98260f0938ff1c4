# quick correctness + speed check of a kernel change: the parity suites that exercise the pixel kernels, then the bench line condensed
python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py tests/test_streams_gpu.py tests/test_overflow_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -${1:-8}
python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-strong 2>gpurun_out/quick_bench.err | tee gpurun_out/quick_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_frame']
print('value %.0f e2e %.0f host %.3f | ' % (d['value'], d['e2e']['value'], d['host_submit_ms_per_step']) + ' '.join('%s=%.0f' % (n.replace('sgl','').replace('Kernel',''), t*1e3) for n,t in sorted(k.items())))"
tail -3 gpurun_out/quick_bench.err
