mkdir -p gpurun_out
python tools/time_variants.py $(cat tools/gpu/tv_list.txt) > gpurun_out/tv.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_selection.py tests/test_gpu_production.py -m gpu -q 2>&1 | tail -3 > gpurun_out/tv_pytest.log
