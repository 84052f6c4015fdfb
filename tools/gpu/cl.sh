mkdir -p gpurun_out
timeout 300 python tools/dbg_cluster_stats.py > gpurun_out/cl_stats.log 2>&1
