set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg_smi.log
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q --tb=short 2>&1 | tail -25 > gpurun_out/mg_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/mg_bench2.log 2>&1
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/mg_smoke.log 2>&1
echo done
