# config 2 bench line condensed, once per environment switch given ("X=0" = defaults)
for sw in "X=0" "$@" "X=0"; do
env $sw python bench.py --steps 500 --warmup 50 --no-cpu-baseline --no-strong 2>gpurun_out/quick_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_frame']
print('$sw c2 value %.0f e2e %.0f host %.3f | ' % (d['value'], d['e2e']['value'], d['host_submit_ms_per_step']) + ' '.join('%s=%.0f' % (n.replace('sgl','').replace('Kernel',''), t*1e3) for n,t in sorted(k.items())))"
done
