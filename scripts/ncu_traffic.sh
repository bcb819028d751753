#!/bin/bash
# DRAM traffic of the dominant kernels of one keyless proof, for bench.py's roofline.traffic (never a bench value):
#   gpurun -- scripts/ncu_traffic.sh            -> gpurun_out/r02_traffic_raw.csv, gpurun_out/r02_traffic.json
# then copy gpurun_out/r02_traffic.json to profiles/ (bench.py reads profiles/r02_traffic.json).
set -e
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --csv --page raw --log-file gpurun_out/r02_traffic_raw.csv \
    -k regex:'k_msm_accumulate|k_ntt_level|k_ntt_mid|k_spmv_abc|k_h_pointwise' \
    python scripts/prove_once.py keyless 2 > gpurun_out/r02_traffic_prove.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r02_traffic_raw.csv gpurun_out/r02_traffic.json
