set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_binning.py tests/test_gpu_production.py -m gpu -q -s -k "ks_statistic or per_sample_values or capped" 2>&1 | grep -v Warning | tail -150 > gpurun_out/c4_pytest.log
echo done
