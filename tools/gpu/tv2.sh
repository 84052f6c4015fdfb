mkdir -p gpurun_out
: > gpurun_out/tv2.log
timeout 300 python tools/time_wide.py >> gpurun_out/tv2.log 2>&1
MCZ_B200_LIB=dbgbuild/libmcz_pt.so timeout 100 python tools/dbg_phases_wide.py >> gpurun_out/tv2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_wide.py -m gpu -q -x --tb=short --show-capture=no 2>&1 | tail -3 >> gpurun_out/tv2.log
