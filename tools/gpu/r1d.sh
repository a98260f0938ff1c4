mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_configs_gpu.py -x -q -m gpu -s 2>&1 | tail -25
timeout 600 python tools/bench_configs.py --cpu --out gpurun_out/r1d_configs.json 2>&1 | tail -8
