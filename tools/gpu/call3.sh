set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | tail -120 > gpurun_out/c3_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c3_smoke.log 2>&1
echo done
