set -u
R=r02
for w in c3; do
  python tools/prof_c2.py $w > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on \
      -k regex:"block_ilu0_lower|block_ilu0_upper|block5_|tri5_|tri_block|bsr|csr_stream|csr_spmv|scalar_lower|scalar_upper" \
      -s 12 -c 14 -o /tmp/prof_${w}_$R -f python tools/prof_c2.py $w > gpurun_out/ncu_$w.log 2>&1
  tail -n 1 gpurun_out/ncu_$w.log
  ncu -i /tmp/prof_${w}_$R.ncu-rep --page raw --csv > gpurun_out/raw_${w}_$R.csv 2>/dev/null
  grep algorithmic_bytes gpurun_out/plain_$w.log > gpurun_out/algbytes_${w}_$R.txt
done
python -c "import __graft_entry__ as g; g.smoke()"
