mkdir -p gpurun_out
python tools/time_variants.py default > gpurun_out/tv.log 2>&1
timeout 2400 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_wide.py tests/test_gpu_replay.py -m gpu -q --tb=short --show-capture=no 2>&1 | grep -v Warn | tail -40 > gpurun_out/tv_pytest.log
