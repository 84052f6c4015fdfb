mkdir -p gpurun_out
export MCZ_CLUSTER=0
timeout 200 python tools/time_variants.py --g 5000 --n 11000 default dbgbuild/libmcz_bu1.so dbgbuild/libmcz_bu3.so dbgbuild/libmcz_bu4.so > gpurun_out/bu_tv.log 2>&1
timeout 200 python tools/time_variants.py --g 20000 --n 11000 --md KD02 default dbgbuild/libmcz_bu1.so dbgbuild/libmcz_bu3.so dbgbuild/libmcz_bu4.so >> gpurun_out/bu_tv.log 2>&1
