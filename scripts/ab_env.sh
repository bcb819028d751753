#!/bin/bash
# A/B runs of bench.py under different environment switches: scripts/ab_env.sh out_prefix "VAR=1 VAR2=x" "VAR=2" ...
# prints ms_per_step / e2e p50 / stage medians / accumulate launch for each
prefix=$1; shift
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${prefix}_$i.json 2> gpurun_out/${prefix}_$i.err
  python - "$cfg" gpurun_out/${prefix}_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    s=d["stage_ms_median"]
    print("%-40s step %.3f e2e_p50 %.3f | spmv %.2f ntt %.2f msm_h %.2f wsort %.2f wg1 %.2f wg2 %.2f | acc %.3f" % (sys.argv[1], d["ms_per_step"], d["e2e"]["latency_ms_p50"], s["spmv_ms"], s["ntt_ms"], s["msm_h_ms"], s["msm_wsort_ms"], s["msm_wg1_ms"], s["msm_wg2_ms"], d["roofline"]["launch_ms"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
