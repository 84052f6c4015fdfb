set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg8_smi.log
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg8_bench$N.log 2>&1
timeout 300 $TR --nproc-per-node $N --master-port 29613 tools/time_wide_sharded.py > gpurun_out/mg8_wide$N.log 2>&1
timeout 300 $TR --nproc-per-node 1 --master-port 29614 tools/time_wide_sharded.py > gpurun_out/mg8_wide1.log 2>&1
timeout 400 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/mg8_bench4.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29615 tools/time_wide_sharded.py > gpurun_out/mg8_wide2.log 2>&1
timeout 400 $TR --nproc-per-node $N --master-port 29616 tests/multigpu/check_sample_sharded.py > gpurun_out/mg8_check$N.log 2>&1
echo done
