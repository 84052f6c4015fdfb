set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29615 tools/time_wide_sharded.py > gpurun_out/mg2_wide2.log 2>&1
timeout 400 $TR --nproc-per-node 2 --master-port 29616 tests/multigpu/check_sample_sharded.py > gpurun_out/mg2_check2.log 2>&1
timeout 400 python -m pytest tests/test_gpu_multigpu.py -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/mg2_pytest.log
echo done
