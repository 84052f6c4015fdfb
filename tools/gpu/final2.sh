set -x
mkdir -p gpurun_out/profiles_r2
export MCZ_PROFILE_DIR=gpurun_out/profiles_r2
export MCZ_COMMIT=$(cat tools/gpu/commit.txt)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f2_smoke.log
timeout 1500 python -u -m pytest tests -m gpu -q -x --tb=short --show-capture=no > gpurun_out/f2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f2_pytest.log
timeout 600 python bench.py > gpurun_out/f2_bench4.log 2>&1
for n in 1100000 11000000; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 1 -c 1 -f -o /tmp/prof_w python tools/prof_other.py wide $n > gpurun_out/prof_ncu_w$n.log 2>&1
python tools/ncu_summary.py /tmp/prof_w.ncu-rep 1 r2_wide_$n $n all wide_kernel > /dev/null 2>> gpurun_out/prof_ncu_w$n.log
rm -f /tmp/prof_w.ncu-rep
done
ls gpurun_out/profiles_r2
echo done
