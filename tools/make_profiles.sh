#!/bin/bash
# Run under gpurun: produces the ncu launch list of the bench command and full captures of the top
# kernels of every BASELINE config at its full size.  Reports are kept outside gpurun_out/ (64 MiB
# cap); only CSV exports travel back.  tools/summarise_profiles.py turns them into profiles/.
set -u
R=${1:-r02}
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --fgmres-n 0"
$BENCH > gpurun_out/bench_plain_$R.log 2> gpurun_out/bench_plain_$R.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$R.csv $BENCH > gpurun_out/ncu_launches_$R.log 2>&1
for w in c2 c3 c1 c4; do
  python tools/prof_c2.py $w > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on \
      -k regex:"block_ilu0_lower|block_ilu0_upper|block5_|tri5_|tri_block|bsr|csr_stream|csr_spmv|scalar_lower|scalar_upper" \
      -s 12 -c 14 -o /tmp/prof_${w}_$R -f python tools/prof_c2.py $w > gpurun_out/ncu_$w.log 2>&1
  tail -n 1 gpurun_out/ncu_$w.log
  ncu -i /tmp/prof_${w}_$R.ncu-rep --page raw --csv > gpurun_out/raw_${w}_$R.csv 2>/dev/null
  grep algorithmic_bytes gpurun_out/plain_$w.log > gpurun_out/algbytes_${w}_$R.txt
done
ls -la gpurun_out | tail -n 20
