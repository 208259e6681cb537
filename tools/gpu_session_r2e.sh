#!/bin/bash
# Round-2 GPU session E (8 GPUs): launch-setting sweep of the REAL N=8 and N=4 syncs inside one job each.
set -u
OUT=gpurun_out/r2e
mkdir -p $OUT
SWEEP="TSB_LINK=1;TSB_LINK=0;TSB_LINK_STAGES=6,TSB_LINK_STAGE_BYTES=8192;TSB_LINK_STAGES=6;TSB_LINK_STAGE_BYTES=8192;TSB_LINK_STAGES=4,TSB_LINK_STAGE_BYTES=2048;TSB_LINK_STAGES=8,TSB_LINK_STAGE_BYTES=16384;TSB_CTAS_PER_SM=2;TSB_CTAS_PER_SM=2,TSB_LINK_STAGES=6;TSB_CTAS_PER_SM=4;TSB_TILE_BYTES=32768;TSB_TILE_BYTES=131072;TSB_COPY_UNROLL=2;TSB_LINK=0,TSB_CTAS_PER_SM=2;TSB_LINK=0,TSB_TILE_BYTES=32768;TSB_LINK=0,TSB_CTAS_PER_SM=4;TSB_LINK=1"
for n in 8 4; do
  echo "== sweep real N=$n"
  TSB_BENCH_SWEEP="$SWEEP" TSB_BENCH_DEADLINE_S=600 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_sweep_n$n.json 2> $OUT/bench_sweep_n$n.err
  echo "rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_sweep_n$n.json"))
print("headline kernel", d["roofline"]["kernel_ms_avg"], "ms/step", d["ms_per_step"])
for r in sorted(d["config"]["launch_setting_sweep"], key=lambda r:r["kernel_ms_avg_max_rank"]): print("%.4f  min %.4f  wall %.4f  %s" % (r["kernel_ms_avg_max_rank"], r["kernel_ms_min_over_steps_max_rank"], r["wall_ms_per_step_max_rank"], r["env"]))
PY
done
