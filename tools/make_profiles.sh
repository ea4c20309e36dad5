#!/bin/bash
# Run under gpurun: produces the ncu launch list of the bench command and full captures of the top
# kernels.  Reports are kept outside gpurun_out/ (64 MiB cap); only CSV exports and one small
# report of the dominant kernel travel back.  tools/summarise_profiles.py turns them into profiles/.
set -u
R=${1:-r01}
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --fgmres-n 0"
$BENCH > gpurun_out/bench_plain_$R.log 2> gpurun_out/bench_plain_$R.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$R.csv $BENCH > gpurun_out/ncu_launches_$R.log 2>&1
for w in c2 c3s p128; do
  python tools/prof_c2.py $w > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on \
      -k regex:"block_ilu0_lower|block_ilu0_upper|tri_block|bsr_spmv|csr_stream|scalar_lower|scalar_upper" \
      -c 16 -o /tmp/prof_${w}_$R -f python tools/prof_c2.py $w > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log
  ncu -i /tmp/prof_${w}_$R.ncu-rep --page raw --csv > gpurun_out/raw_${w}_$R.csv 2>/dev/null
done
# one small report of the dominant kernel (C2 upper launch) with source correlation
ncu --set full --clock-control none --import-source on -k regex:"block_ilu0_upper" -s 1 -c 1 \
    -o gpurun_out/prof_c2_upper_$R -f python tools/prof_c2.py c2 > gpurun_out/ncu_c2_upper.log 2>&1
ls -la gpurun_out | tail -20
