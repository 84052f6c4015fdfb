mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -m gpu -q --tb=short --show-capture=no 2>&1 | grep -v Warn | tail -15 > gpurun_out/wd_pytest.log
timeout 600 python tools/time_wide.py > gpurun_out/wd_time.log 2>&1
