set -u
python bench.py --steps 20 --warmup 5 --fgmres-n 0 --no-cpu > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/c7_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), round(v['achieved_gbs']), round(v['share_of_step'],3))
PY
python tools/ab_factor.py c2 > gpurun_out/c7_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:"tri_block_pipe" -s 8 -c 4 -f -o gpurun_out/prof_c2_tri_r02a python tools/ab_factor.py c2 > gpurun_out/c7_ncu.log 2>&1
tail -n 2 gpurun_out/c7_plain.log
