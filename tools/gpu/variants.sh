# A/B of compile-time variants built with `python -m softglrender_b200.build --variant NAME ...`: condensed config-2 bench line each
for v in "" $(ls softglrender_b200/lib_variants 2>/dev/null) ""; do
  if [ -n "$v" ]; then export SGL_LIB_DIR=$PWD/softglrender_b200/lib_variants/$v; else unset SGL_LIB_DIR; fi
  python bench.py --steps 400 --warmup 40 --no-cpu-baseline --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_frame']
print('%-8s value %.0f e2e %.0f host %.3f | ' % ('${v:-base}', d['value'], d['e2e']['value'], d['host_submit_ms_per_step']) + ' '.join('%s=%.0f' % (n.replace('sgl','').replace('Kernel',''), t*1e3) for n,t in sorted(k.items())))"
done
