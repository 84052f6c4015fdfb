mkdir -p gpurun_out/profiles_r2
export MCZ_PROFILE_DIR=gpurun_out/profiles_r2
export MCZ_COMMIT=$(cat tools/gpu/commit.txt)
n=11000000
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 1 -c 1 -f -o /tmp/prof_w python tools/prof_other.py wide $n > gpurun_out/prof_ncu_w2_$n.log 2>&1
python tools/ncu_summary.py /tmp/prof_w.ncu-rep 1 r2b_wide_$n $n all wide_kernel > /dev/null 2>> gpurun_out/prof_ncu_w2_$n.log
