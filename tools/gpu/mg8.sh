set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $N --master-port 29613 tools/time_wide_sharded.py > gpurun_out/mg8b_wide$N.log 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29614 tools/time_wide_sharded.py > gpurun_out/mg8b_wide4.log 2>&1
timeout 400 $TR --nproc-per-node $N --master-port 29616 tests/multigpu/check_sample_sharded.py > gpurun_out/mg8b_check$N.log 2>&1
timeout 600 $TR --nproc-per-node $N --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/mg8b_bench$N.log 2>&1
echo done
