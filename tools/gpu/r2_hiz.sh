python -m pytest tests/test_parity_configs_gpu.py tests/test_parity_gpu.py tests/test_multigpu_gpu.py tests/test_overflow_gpu.py tests/test_streams_gpu.py -m gpu -q -x 2>&1 | tail -8
python tools/bench_configs.py --only c4,c4big,c4full,c3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l[:1]=='c' and ' {' in l:
        k,_,j=l.partition(' '); d=json.loads(j); print(k, round(d['units_per_s'],1), 'fps', d['bin_entries_per_step'], {a: round(b,2) for a,b in d['kernel_ms_per_step'].items()}, d.get('roofline',{}).get('frac'))"
bash tools/gpu/quick.sh 1 2>&1 | tail -2
