#!/bin/bash
# A/B runs of tools/group_bench.py under different environment switches:
#   scripts/ab_group.sh devices out_prefix "VAR=1 VAR2=x" "VAR=2" ...
devs=$1; prefix=$2; shift; shift
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python tools/group_bench.py --devices $devs --no-check --steps 12 --out gpurun_out/${prefix}_$i.json > /dev/null 2> gpurun_out/${prefix}_$i.err
  python - "$cfg" gpurun_out/${prefix}_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    s=d["shard_stage_ms_median"]
    mx=lambda k: max(x[k] for x in s)
    print("%-44s res p50 %.3f min %.3f e2e %.3f | spmv %.2f ntt %.2f msm_h %.2f wsort %.2f wg1 %.2f wg2 %.2f gpu %.2f" % (sys.argv[1], d["resident_ms_p50"], d["resident_ms_min"], d["e2e_ms_p50"], mx("spmv_ms"), mx("ntt_ms"), mx("msm_h_ms"), mx("msm_wsort_ms"), mx("msm_wg1_ms"), mx("msm_wg2_ms"), mx("gpu_ms")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
