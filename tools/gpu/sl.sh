mkdir -p gpurun_out
MCZ_B200_LIB=dbgbuild/libmcz_ctA.so timeout 200 python tools/dbg_slowcols_long.py 2>&1 | grep "crowded" > gpurun_out/sl.log
MCZ_B200_LIB=dbgbuild/libmcz_ctB.so timeout 200 python tools/dbg_slowcols_long.py 2>&1 | grep "crowded" >> gpurun_out/sl.log
