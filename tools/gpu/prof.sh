set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2c python tools/prof_one.py > gpurun_out/prof_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
