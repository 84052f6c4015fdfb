set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f3_smoke.log
timeout 1500 python -u -m pytest tests -m gpu -q -x --tb=short --show-capture=no > gpurun_out/f3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f3_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/f3_bench4.log 2>&1
echo done
