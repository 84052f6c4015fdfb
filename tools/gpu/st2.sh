mkdir -p gpurun_out
: > gpurun_out/st2.log
for md in KD02 all; do for pr in 0 1; do
echo "== md $md MCZ_PAIR=$pr" >> gpurun_out/st2.log
MCZ_PAIR=$pr MCZ_B200_LIB=dbgbuild/libmcz_ct2.so timeout 100 python tools/dbg_stages.py ct2 11000 1480 $md >> gpurun_out/st2.log 2>&1
MCZ_PAIR=$pr MCZ_B200_LIB=dbgbuild/libmcz_ct3.so timeout 100 python tools/dbg_stages.py ct3 11000 1480 $md >> gpurun_out/st2.log 2>&1
done; done
