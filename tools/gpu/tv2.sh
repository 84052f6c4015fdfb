mkdir -p gpurun_out
: > gpurun_out/tv2.log
MCZ_B200_LIB=dbgbuild/libmcz_ct2.so timeout 200 python tools/dbg_slowcols_long.py 2>&1 | grep "crowded" >> gpurun_out/tv2.log
timeout 300 python tools/time_variants.py --g 5000 --n 11000 default dbgbuild/libmcz_base.so >> gpurun_out/tv2.log 2>&1
timeout 300 python tools/time_variants.py --g 20000 --n 11000 --md KD02 default dbgbuild/libmcz_base.so >> gpurun_out/tv2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_selection.py tests/test_gpu_production.py -m gpu -q -x --tb=short --show-capture=no 2>&1 | tail -3 >> gpurun_out/tv2.log
