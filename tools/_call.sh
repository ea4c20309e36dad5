set -u
timeout 900 python -m pytest tests -m gpu -q -x -k "setup or level or ilu or sgs or edge" > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.log
grep -v "Step" gpurun_out/c11_pytest.log | tail -n 4 | cut -c1-200
timeout 1200 python tools/config_report.py c1 c2 c3 c4 frontend > gpurun_out/configs_r02.md 2> gpurun_out/c11_cfg.err; echo "config rc=$?"
grep -v "Step" gpurun_out/configs_r02.md | grep -i "level\|compute()\|FGMRES\|GCR" | cut -c1-220
tail -n 3 gpurun_out/c11_cfg.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_r02.json 2> gpurun_out/c11_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_r02.json'))
print(d['value'], d['ms_per_step'], json.dumps(d['e2e']))
print(d['fgmres'])
PY
