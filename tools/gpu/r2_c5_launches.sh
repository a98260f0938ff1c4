# cold-cache serialised launch list of config 5 (64 views of AfricanHead 512x512): the true per-kernel GPU times of the small passes
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_c5.csv python tools/bench_configs.py --only c5 > gpurun_out/r02_launches_c5.log 2>&1
python tools/launch_table.py gpurun_out/r02_launches_c5.csv
