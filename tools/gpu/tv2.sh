mkdir -p gpurun_out
: > gpurun_out/tv2.log
timeout 300 python tools/time_variants.py --g 5000 --n 11000 default dbgbuild/libmcz_gen3.so default dbgbuild/libmcz_gen3.so >> gpurun_out/tv2.log 2>&1
timeout 300 python tools/time_variants.py default dbgbuild/libmcz_gen3.so >> gpurun_out/tv2.log 2>&1
MCZ_B200_LIB=dbgbuild/libmcz_gen3.so timeout 900 python -m pytest tests/test_gpu_selection.py tests/test_gpu_production.py -m gpu -q -x --tb=short --show-capture=no 2>&1 | tail -3 >> gpurun_out/tv2.log
