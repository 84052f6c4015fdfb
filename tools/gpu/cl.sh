mkdir -p gpurun_out
export MCZ_CLUSTER=1
export MCZ_DEBUG_POISON=1
timeout 100 python tools/time_variants.py --g 2000 --n 11000 default > gpurun_out/cl_tv.log 2>&1 || { echo "first timing failed rc=$?" >> gpurun_out/cl_tv.log; exit 1; }
timeout 100 python tools/time_variants.py --g 5000 --n 11000 default >> gpurun_out/cl_tv.log 2>&1
timeout 100 python tools/time_variants.py --g 20000 --n 11000 --md KD02 default >> gpurun_out/cl_tv.log 2>&1
timeout 100 python tools/time_variants.py --g 10000 --n 4608 default >> gpurun_out/cl_tv.log 2>&1
MCZ_CLUSTER_SIZE=6 timeout 100 python tools/time_variants.py --g 5000 --n 11000 default >> gpurun_out/cl_tv.log 2>&1
MCZ_CLUSTER_SIZE=7 timeout 100 python tools/time_variants.py --g 5000 --n 11000 default >> gpurun_out/cl_tv.log 2>&1
MCZ_CLUSTER=0 timeout 100 python tools/time_variants.py --g 5000 --n 11000 default >> gpurun_out/cl_tv.log 2>&1
timeout 400 python -u -m pytest tests/test_gpu_production.py -m gpu -q -x --tb=short --show-capture=no -k "cluster" > gpurun_out/cl_pytest.log 2>&1
