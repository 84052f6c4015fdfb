mkdir -p gpurun_out
export MCZ_CLUSTER=0
MCZ_B200_LIB=dbgbuild/libmcz_ct2.so timeout 100 python tools/dbg_stages.py "rangeH,plan,G" 11000 1480 > gpurun_out/st.log 2>&1
MCZ_B200_LIB=dbgbuild/libmcz_ct3.so timeout 100 python tools/dbg_stages.py "range,finish,total" 11000 1480 >> gpurun_out/st.log 2>&1
