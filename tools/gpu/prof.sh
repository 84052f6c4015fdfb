set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_wide python tools/prof_other.py wide > gpurun_out/prof_ncu_w.log 2>&1
MCZ_CLUSTER=1 ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_cluster python tools/prof_other.py global > gpurun_out/prof_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
