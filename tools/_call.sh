set -u
timeout 900 python -m pytest tests -m gpu -x -q -k "ilu or full or edge" > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18_pytest.log
grep -v "Step" gpurun_out/c18_pytest.log | tail -n 3 | cut -c1-220
for i in 1 2; do
python tools/ab_factor.py c3 2>&1 | tail -n 1
B200_NO_STAGED_LOWER=1 python tools/ab_factor.py c3 2>&1 | tail -n 1
done
