mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29616 tests/multigpu/check_sample_sharded.py > gpurun_out/mg2_check2.log 2>&1
echo "rc=$?" >> gpurun_out/mg2_check2.log
