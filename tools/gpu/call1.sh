set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_smi.log 2>&1
nproc >> gpurun_out/c1_smi.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/c1_pytest.log
timeout 900 python tools/trip_parity.py 2000 2200 all fx20 > gpurun_out/c1_trip_fx20.log 2>&1
MCZ_B200_LIB=$PWD/dbgbuild/libmcz_fx16.so timeout 900 python tools/trip_parity.py 2000 2200 all fx16 > gpurun_out/c1_trip_fx16.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench4.log 2>&1
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench3.log 2>&1
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench5.log 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/c1_benchref.log 2>&1
echo done
