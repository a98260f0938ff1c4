# Round profile set (single GPU): ncu --set full of the two dominant kernels, ncu launch list of a short bench run,
# the full bench line (with cpu_baseline) and the reference arm.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --target-processes application-only -k regex:"sglShadeKernel|sglVisKernel" -s 8 -c 2 \
    -o gpurun_out/${tag}_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --target-processes application-only -s 80 -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/${tag}_ncu_launches.log 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
tail -c 700 gpurun_out/${tag}_bench.json; tail -c 400 gpurun_out/${tag}_bench_reference.json
