mkdir -p gpurun_out
MCZ_B200_LIB=dbgbuild/libmcz_ct3.so timeout 200 python tools/dbg_slowcols_long.py > gpurun_out/sl.log 2>&1
