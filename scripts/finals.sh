set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 600 gpurun_out/r2f_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_bench_under_ncu.log 2>&1
bash scripts/ncu_traffic.sh
python tools/microbench.py --g2 16,18,20,21,22,24 --out gpurun_out/r2f_microbench.json > gpurun_out/r2f_microbench.log 2>&1
tail -2 gpurun_out/r2f_microbench.log
bash scripts/ncu_full_proof.sh
