set -u
timeout 600 python -m pytest tests -m gpu -x -q -k "ilu or full" > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
grep -v "Step" gpurun_out/c13_pytest.log | tail -n 3 | cut -c1-200
for i in 1 2; do
python tools/ab_factor.py c3 2>&1 | tail -n 1
B200_NO_RANK1=1 python tools/ab_factor.py c3 2>&1 | tail -n 1
B200_LIB=$PWD/blasted_b200/libblasted_b200_v3.so python tools/ab_factor.py c3 2>&1 | tail -n 1
done
