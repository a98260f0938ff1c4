SGL_LIB_DIR=$PWD/softglrender_b200/lib_variants/touch python tools/gpu/texel_touch.py c2 > gpurun_out/r02_texel_touch_c2.json 2> gpurun_out/touch.err; cat gpurun_out/r02_texel_touch_c2.json; tail -3 gpurun_out/touch.err
python bench.py --steps 200 --warmup 20 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 3000 gpurun_out/r2_bench2.json; tail -3 gpurun_out/r2_bench2.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench2_ref.json 2>gpurun_out/r2_bench2_ref.err; tail -c 600 gpurun_out/r2_bench2_ref.json
