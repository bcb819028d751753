#!/bin/bash
# The measurements DESIGN.md / profiles/README.md quote for the final code, one gpurun call:
#   gpurun --timeout 900 -- bash scripts/finals.sh     (then copy gpurun_out/r2f_* and r02_* into profiles/)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
tail -2 gpurun_out/r2f_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 300 gpurun_out/r2f_bench.json
python bench.py --workload keyless-wide --no-cpu-baseline > gpurun_out/r2f_bench_wide.json 2> gpurun_out/r2f_bench_wide.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_bench_under_ncu.log 2>&1
bash scripts/ncu_traffic.sh
bash scripts/ncu_full_proof.sh
