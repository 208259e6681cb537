#!/bin/bash
# Round-2 GPU session G (8 GPUs): final-tree confirmation of the headline (fan-in ring depth), init phases,
# context listing, all-remote configs with the deeper ring, store sweep with the epoch board.
set -u
OUT=gpurun_out/r2g
mkdir -p $OUT
run_bench() { n=$1; name=$2; shift; shift
  echo "== bench $name (N=$n)"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "rc=$?"; head -c 300 $OUT/bench_$name.json; echo; grep -E "PARITY|Error|error|Traceback|deadline" $OUT/bench_$name.err | head -5
}
SWEEP="TSB_LINK_STAGES=8;TSB_LINK_STAGES=5;TSB_LINK_STAGES=4;TSB_LINK_STAGES=3;TSB_LINK_STAGES=6,TSB_LINK_STAGE_BYTES=2048;TSB_LINK_STAGES=8,TSB_LINK_STAGE_BYTES=2048;TSB_LINK=0"
TSB_BENCH_SWEEP="$SWEEP" TSB_BENCH_LIST_CONTEXTS=1 TSB_TRACE_INIT=1 run_bench 8 n8 --steps 30 --warmup 3
run_bench 8 cfg3b --config 3b --steps 10 --warmup 3
run_bench 8 cfg3a --config 3a --steps 10 --warmup 3
run_bench 8 cfg5 --config 5
python - <<PY
import json
d=json.load(open("$OUT/bench_n8.json"))
print("n8 value %.0f ms/step %.4f kernel %.4f frac %.3f e2e %.1f overhead %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["e2e"]["value"], d["config"]["host_overhead_ms_per_step"]))
for r in sorted(d["config"].get("launch_setting_sweep", []), key=lambda r:r["kernel_ms_avg_max_rank"]): print("  %.4f  %s" % (r["kernel_ms_avg_max_rank"], r["env"]))
for f in ("cfg3a","cfg3b"):
    d=json.load(open("$OUT/bench_%s.json"%f)); print(f, "value %.0f kernel %.3f frac %.3f" % (d["value"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"]))
d=json.load(open("$OUT/bench_cfg5.json"))
for r in d["config"]["sweep"]: print("cfg5", r["key_bytes"]>>20, "MiB x", r["keys_per_batch"], "put %.0f get %.0f peer %s GB/s/GPU" % (r["put_GBps_per_gpu"], r["get_local_GBps_per_gpu"], r["get_peer_GBps_per_gpu"]))
PY
grep -E "r0 |spmd r0|compute contexts" -A0 $OUT/bench_n8.err | head -40
grep -A 20 "compute contexts" $OUT/bench_n8.err | head -24
