mkdir -p gpurun_out
export MCZ_CLUSTER=0
timeout 200 python tools/time_variants.py --g 5000 --n 11000 default > gpurun_out/gl_tv.log 2>&1 || { echo "first timing failed rc=$?" >> gpurun_out/gl_tv.log; exit 1; }
timeout 200 python tools/time_variants.py --g 20000 --n 11000 --md KD02 default >> gpurun_out/gl_tv.log 2>&1
timeout 200 python tools/time_variants.py --g 10000 --n 4608 default >> gpurun_out/gl_tv.log 2>&1
timeout 900 python -u -m pytest tests/test_gpu_production.py -m gpu -q -x --tb=short --show-capture=no -k "large_nsample or cluster or capped" > gpurun_out/gl_pytest.log 2>&1
