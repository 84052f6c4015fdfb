set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/c2_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c2_bench4.log 2>&1
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench3.log 2>&1
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench5.log 2>&1
echo done
