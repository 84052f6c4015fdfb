set -x
mkdir -p gpurun_out
# launch list of the default bench command (after it has exited 0 without ncu)
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_r2.log 2>&1
# one ncu --set full capture per bench launch shape (second launch of each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_cfg4 python tools/prof_bench_shape.py 4 > gpurun_out/prof_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_cfg3 python tools/prof_bench_shape.py 3 > gpurun_out/prof_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_cfg5 python tools/prof_bench_shape.py 5 > gpurun_out/prof_ncu5.log 2>&1
MCZ_CLUSTER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_cluster python tools/prof_bench_shape.py 3 2000 > gpurun_out/prof_ncuc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_wide python tools/prof_other.py wide > gpurun_out/prof_ncuw.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
