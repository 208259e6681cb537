#!/bin/bash
# Round-2 GPU session B (2 GPUs): racecheck harness, in-flight-depth sweep of the emulated N=8/4/2 syncs,
# and a first pass of the new bench.py (all configs at N=2).
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== racecheck harness"
timeout 300 compute-sanitizer --tool racecheck --racecheck-report all experiments/racecheck_harness.bin > $OUT/racecheck_harness.log 2>&1; echo "rc=$?" >> $OUT/racecheck_harness.log; tail -6 $OUT/racecheck_harness.log
timeout 300 compute-sanitizer --tool memcheck experiments/racecheck_harness.bin > $OUT/memcheck_harness.log 2>&1; echo "rc=$?" >> $OUT/memcheck_harness.log; tail -4 $OUT/memcheck_harness.log
timeout 300 compute-sanitizer --tool synccheck experiments/racecheck_harness.bin > $OUT/synccheck_harness.log 2>&1; echo "rc=$?" >> $OUT/synccheck_harness.log; tail -4 $OUT/synccheck_harness.log
echo "== link tests"; timeout 300 python -m pytest tests/test_gpu_copy_rects.py -x -q > $OUT/pytest_copy_rects.log 2>&1; tail -3 $OUT/pytest_copy_rects.log
for N in 8 4 2; do
  echo "== sweep x2 n=$N"
  timeout 900 python tools/sweep_plan.py --mode x2 --n $N --iters 4 --env "$(python tools/sweep_x2_grid.py $N)" --out $OUT/sweep_x2_n$N.json > $OUT/sweep_x2_n$N.log 2>&1
  python - <<PY
import json
rows=json.load(open("$OUT/sweep_x2_n$N.json"))
rows.sort(key=lambda r:r["ms_median"])
for r in rows[:8]: print(r["ms_median"], r["env"])
print("...worst", rows[-1]["ms_median"], rows[-1]["env"])
PY
done
run_bench() { # name, extra args
  name=$1; shift
  echo "== bench $name"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "rc=$?"; head -c 1500 $OUT/bench_$name.json; echo; grep -E "PARITY|Error|error|Traceback" $OUT/bench_$name.err | head -5
}
run_bench n2 --steps 20 --warmup 3
run_bench cfg2 --config 2 --steps 10 --warmup 3
run_bench cfg4p --config 4p --steps 10 --warmup 3
run_bench cfg3a --config 3a --steps 10 --warmup 3
run_bench cfg3b --config 3b --steps 10 --warmup 3
run_bench cfg3b_nccl --config 3b --allgather --steps 10 --warmup 3
TSB_SWEEP_BYTES=4294967296 run_bench cfg5 --config 5
echo "== bench n1"; timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; head -c 1500 $OUT/bench_n1.json
ls -la $OUT
