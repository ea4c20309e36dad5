set -u
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c27_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c27_pytest.log
grep -v "Step" gpurun_out/c27_pytest.log | tail -n 4 | cut -c1-220
timeout 600 python tools/config_report.py c1 2>/dev/null | grep -v Step | grep -i "sweep (apply)\|SGS apply\|apply()\|FGMRES" | cut -c1-200
B200_NO_PERSIST_STREAM=1 timeout 600 python tools/config_report.py c1 2>/dev/null | grep -v Step | grep -i "sweep (apply)\|SGS apply\|apply()\|FGMRES" | cut -c1-200
