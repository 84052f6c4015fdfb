mkdir -p gpurun_out
export MCZ_CLUSTER=1
timeout 300 python tools/time_variants.py --g 5000 --n 11000 default > gpurun_out/cl_tv.log 2>&1
timeout 300 python tools/time_variants.py --g 20000 --n 11000 --md KD02 default >> gpurun_out/cl_tv.log 2>&1
timeout 300 python tools/time_variants.py --g 10000 --n 4608 default >> gpurun_out/cl_tv.log 2>&1
timeout 600 python -m pytest tests/test_gpu_production.py -m gpu -q -x --tb=short --show-capture=no -k "cluster or large_nsample" 2>&1 | grep -v Warn | tail -12 > gpurun_out/cl_pytest.log
