#!/bin/bash
# Round-2 GPU session I (2 GPUs): EXPERIMENT -- L2 staging of served memory (UBLKPF.L2 evict-last, paced by the link
# queue) and evict-first policy on the copy stream, on the 2-GPU emulation of the N=8/4/2 syncs.
set -u
OUT=gpurun_out/r2i
mkdir -p $OUT
ENVS=";TSB_L2_STREAM=1;TSB_STAGE=1;TSB_STAGE=1,TSB_L2_STREAM=1;TSB_STAGE=1,TSB_L2_STREAM=1,TSB_STAGE_LEAD_BYTES=1048576;TSB_STAGE=1,TSB_L2_STREAM=1,TSB_STAGE_LEAD_BYTES=2097152;TSB_STAGE=1,TSB_L2_STREAM=1,TSB_STAGE_LEAD_BYTES=4194304;TSB_STAGE=1,TSB_L2_STREAM=1,TSB_STAGE_LEAD_BYTES=16777216;TSB_STAGE=1,TSB_L2_STREAM=1,TSB_STAGE_LEAD_BYTES=33554432;TSB_STAGE=1,TSB_STAGE_LEAD_BYTES=2097152;TSB_STAGE=1,TSB_STAGE_LEAD_BYTES=33554432;TSB_STAGE=1,TSB_L2_STREAM=1,TSB_STAGE_LEAD_BYTES=262144;TSB_STAGE=1,TSB_L2_STREAM=1,TSB_STAGE_LEAD_BYTES=0"
for N in 8 4 2; do
  echo "== x2 n=$N"
  timeout 600 python tools/sweep_plan.py --mode x2 --n $N --iters 5 --env "$ENVS" --out $OUT/stage_x2_n$N.json > $OUT/stage_x2_n$N.log 2>&1
  tail -3 $OUT/stage_x2_n$N.log | cut -c1-300
  python - <<PY
import json
try:
    rows=json.load(open("$OUT/stage_x2_n$N.json"))
    for r in sorted(rows,key=lambda r:r["ms_median"]): print("%.4f  nvl %.0f  %s" % (r["ms_median"], r["nvlink_in_GBps"], r["env"] or "(defaults)"))
except Exception as e: print("ERR", e)
PY
done
