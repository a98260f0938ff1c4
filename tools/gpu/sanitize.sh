# compute-sanitizer over the trace player on small traces (KAT with MSAA + reversed-Z, KAT 1x, Cube 256x192), frame section
# replayed three more times (early visibility, renamed shadow map); tools: $1 (default "memcheck racecheck")
mkdir -p gpurun_out build/san
python - > build/san/traces.txt <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests/golden")
import make_golden
from softglrender_b200 import workloads
for n in ("kat_ms4_revz", "kat_1x"):
    make_golden.build_trace(n, "build/san")
    print("build/san/%s.sglt" % n)
print(workloads.build_c1("build/san", 256, 192)[0])
PY
for tool in ${1:-memcheck racecheck}; do
  for t in $(cat build/san/traces.txt); do
    b=$(basename $t .sglt)
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 softglrender_b200/lib/sgl_player $t --data-dir build/san --out build/san/$b.out --frames 3 > gpurun_out/san_${tool}_$b.log 2>&1
    echo "$tool $b rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}_$b.log | tail -1)"
  done
done
