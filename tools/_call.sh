set -u
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
grep -v "Step" gpurun_out/c10_pytest.log | tail -n 60 | cut -c1-260
