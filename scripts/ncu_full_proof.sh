#!/bin/bash
# ncu --set full of every kernel of ONE keyless proof (the second proof of the process: the first one pages the key in):
#   gpurun -- scripts/ncu_full_proof.sh  -> gpurun_out/r02_ncu_full_raw.csv, gpurun_out/r02_ncu_full.txt
set -e
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --launch-skip 37 -c 37 \
    -k regex:'k_spmv|k_ntt|k_h_pointwise|k_msm|k_s2|k_witness' -o /tmp/r02_full python scripts/prove_once.py keyless 2 > gpurun_out/r02_ncu_full_prove.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv
python scripts/ncu_summary.py gpurun_out/r02_ncu_full_raw.csv > gpurun_out/r02_ncu_full.txt
tail -2 gpurun_out/r02_ncu_full_prove.log
