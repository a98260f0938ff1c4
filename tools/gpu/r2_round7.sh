# multisample colour mask: streams suite (switch matrix), parity suites, config 2 with the switch on/off
mkdir -p gpurun_out
python -m pytest tests/test_streams_gpu.py -m gpu -q -x 2>&1 | tail -5
python -m pytest tests/test_parity_gpu.py tests/test_parity_configs_gpu.py tests/test_overflow_gpu.py tests/test_multigpu_gpu.py tests/test_viewer_integration_gpu.py tests/test_unit_kats_gpu.py -m gpu -q -x 2>&1 | tail -3
bash tools/gpu/r2_c2_ab.sh SGL_NO_MS_MASK=1 2>&1 | head -3
