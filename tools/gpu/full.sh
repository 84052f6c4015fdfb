mkdir -p gpurun_out
timeout 1500 python -u -m pytest tests -m gpu -q -x --tb=short --show-capture=no > gpurun_out/full_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/full_pytest.log
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/full_bench3.log 2>&1
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/full_bench5.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/full_bench4.log 2>&1
