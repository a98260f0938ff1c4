# compute-sanitizer over the trace player on small traces (KAT with MSAA + reversed-Z, KAT 1x, Cube 256x192)
mkdir -p gpurun_out build/san
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests/golden")
import make_golden
from softglrender_b200 import workloads
for n in ("kat_ms4_revz", "kat_1x"):
    make_golden.build_trace(n, "build/san")
workloads.build_c1("build/san", 256, 192)
PY
for tool in memcheck racecheck; do
  for t in kat_ms4_revz kat_1x c1_256x192; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 softglrender_b200/lib/sgl_player build/san/$t.sglt --data-dir build/san --out build/san/$t.out > gpurun_out/san_${tool}_$t.log 2>&1
    echo "$tool $t rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_$t.log | tail -1)"
  done
done
