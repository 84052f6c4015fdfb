mkdir -p gpurun_out
export MCZ_CLUSTER=0
MCZ_B200_LIB=dbgbuild/libmcz_pt.so timeout 100 python tools/dbg_phases_global.py > gpurun_out/ph2.log 2>&1
MCZ_B200_LIB=dbgbuild/libmcz_pt.so timeout 100 python tools/dbg_phases_global.py KD02 >> gpurun_out/ph2.log 2>&1
