ncu --set full --import-source on --clock-control none -k regex:'k_ntt_level_tma|k_ntt_mid' -c 3 -o /tmp/r2r_ntt python scripts/prove_once.py keyless 1 > gpurun_out/r2r_prove.log 2>&1
ncu -i /tmp/r2r_ntt.ncu-rep --page raw --csv > gpurun_out/r2r_ntt_raw.csv
python scripts/ncu_summary.py gpurun_out/r2r_ntt_raw.csv > gpurun_out/r2r_ntt_summary.txt
ncu -i /tmp/r2r_ntt.ncu-rep --page source --csv --kernel-name regex:k_ntt_mid > gpurun_out/r2r_mid_source.csv 2>/dev/null
tail -5 gpurun_out/r2r_prove.log
