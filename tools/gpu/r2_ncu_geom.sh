# ncu --set full of the geometry kernels of config 4 (2 M triangles, 8K) as rank 0 of 8 on one GPU
ncu --set full --clock-control none --import-source on -k regex:'sglSetupKernel|sglBinFillKernel|sglVaryingKernel|sglVertexKernel|sglTileSortKernel' \
  --launch-skip 30 -c 5 -o gpurun_out/r02_geom -f python tools/bench_configs.py --only ${1:-c4big} --as-rank ${2:-0/8} > gpurun_out/r02_ncu_geom.log 2>&1
tail -n 3 gpurun_out/r02_ncu_geom.log | cut -c1-200
ncu -i gpurun_out/r02_geom.ncu-rep --page raw --csv > gpurun_out/r02_geom_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_geom_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i].split("__")[-1][:40]: r[i] for i in idx})
PY
