# A/B of compile-time variants built with `python -m softglrender_b200.build --variant NAME ...`
for v in "" $(ls softglrender_b200/lib_variants); do
  if [ -n "$v" ]; then export SGL_LIB_DIR=$PWD/softglrender_b200/lib_variants/$v; else unset SGL_LIB_DIR; fi
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_frame']
print('%-10s value %.0f e2e %.0f  vis %.0f shade %.0f' % ('${v:-base}', d['value'], d['e2e']['value'], k['sglVisKernel<4>']*1e3, k['sglShadeKernel<4>']*1e3))"
done
