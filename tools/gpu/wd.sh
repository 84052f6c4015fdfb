mkdir -p gpurun_out
timeout 300 python -u -m pytest tests/test_gpu_wide.py -m gpu -q -x --tb=short --show-capture=no > gpurun_out/wd_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/wd_pytest.log
timeout 100 python tools/time_wide.py > gpurun_out/wd_time.log 2>&1
MCZ_B200_LIB=dbgbuild/libmcz_pt.so timeout 100 python tools/dbg_phases_wide.py 1100000 > gpurun_out/wd_ph.log 2>&1
MCZ_B200_LIB=dbgbuild/libmcz_pt.so timeout 100 python tools/dbg_phases_wide.py 11000000 >> gpurun_out/wd_ph.log 2>&1
