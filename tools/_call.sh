set -u
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c28_pytest.log
grep -v "Step" gpurun_out/c28_pytest.log | tail -n 6 | cut -c1-220
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --fgmres-n 0 > gpurun_out/c28_bench.json 2> gpurun_out/c28_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/c28_bench.json'))
print(d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:330])"
