set -x
mkdir -p gpurun_out/profiles_r2
export MCZ_PROFILE_DIR=gpurun_out/profiles_r2
export MCZ_COMMIT=$(cat tools/gpu/commit.txt)
cap() {  # tag config galaxies n_draw md kernel_base [extra env]
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:production_kernel -s 1 -c 1 -f -o /tmp/prof_$1 python tools/prof_bench_shape.py $2 $3 > gpurun_out/prof_ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep $3 $1 $4 $5 "$6" > /dev/null 2>> gpurun_out/prof_ncu_$1.log
  rm -f /tmp/prof_$1.ncu-rep
}
timeout 1500 python -u -m pytest tests -m gpu -q -x --tb=short --show-capture=no > gpurun_out/full_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/full_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/prof_bench4.log 2>&1 || exit 1
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench3.log 2>&1
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_r2.log 2>&1
cap r2_cfg3 3 10000 11000 all "production_kernel<0,0>"
cap r2_cfg5 5 100000 11000 KD02 "production_kernel<0,0,pair>"
cap r2_cfg4 4 1000000 2200 all "production_kernel<4,0>"
MCZ_B200_LIB=dbgbuild/libmcz_pt.so timeout 100 python tools/dbg_phases_global.py > gpurun_out/ph2.log 2>&1
MCZ_B200_LIB=dbgbuild/libmcz_pt.so timeout 100 python tools/dbg_phases_global.py KD02 >> gpurun_out/ph2.log 2>&1
echo done
