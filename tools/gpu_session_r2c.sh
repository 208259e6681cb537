#!/bin/bash
# Round-2 GPU session C (2 GPUs): full -m gpu suite, narrow ring sweep, ncu traffic captures of the shipped build
# (N=1 bench + 2-GPU emulations with NVLink counters), launch list, store fast lane numbers.
set -u
OUT=gpurun_out/r2c
mkdir -p $OUT
echo "== pytest -m gpu (2 GPUs)"; timeout 1200 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu_2gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu_2gpu.log; tail -6 $OUT/pytest_gpu_2gpu.log
ENVS="TSB_LINK=1;TSB_LINK=0;TSB_LINK_STAGE_BYTES=2048;TSB_LINK_STAGE_BYTES=1024;TSB_LINK_STAGE_BYTES=2048,TSB_LINK_STAGES=4;TSB_LINK_STAGE_BYTES=1024,TSB_LINK_STAGES=6;TSB_CTAS_PER_SM=2;TSB_CTAS_PER_SM=2,TSB_LINK_STAGE_BYTES=2048;TSB_TILE_BYTES=32768;TSB_TILE_BYTES=131072;TSB_CTAS_PER_SM=2,TSB_TILE_BYTES=131072;TSB_CTAS_PER_SM=4"
for N in 8 4 2; do
  echo "== sweep x2 n=$N"
  timeout 600 python tools/sweep_plan.py --mode x2 --n $N --iters 5 --env "$ENVS" --out $OUT/sweep_x2_n$N.json > $OUT/sweep_x2_n$N.log 2>&1
  python - <<PY
import json
rows=json.load(open("$OUT/sweep_x2_n$N.json"))
for r in sorted(rows,key=lambda r:r["ms_median"]): print(r["ms_median"], r["env"] or "(defaults)")
PY
done
NVL=nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum
for N in 8 4 2; do
  echo "== ncu x2 n=$N"
  timeout 600 ncu --set full --metrics $NVL --clock-control none --import-source on --launch-skip 4 -c 1 -k regex:copy_rects \
      -o $OUT/ncu_x2_n$N -f python tools/sweep_plan.py --mode x2 --n $N --iters 1 > $OUT/ncu_x2_n$N.log 2>&1
  ncu -i $OUT/ncu_x2_n$N.ncu-rep --page raw --csv > $OUT/r2_ncu_traffic_x2_n$N.csv 2>/dev/null
  [ $N != 8 ] && rm -f $OUT/ncu_x2_n$N.ncu-rep
  tail -2 $OUT/ncu_x2_n$N.log
done
echo "== ncu n1 (bench.py)"
timeout 900 ncu --set full --metrics $NVL --clock-control none --import-source on --launch-skip 3 -c 1 -k regex:copy_rects \
    -o $OUT/ncu_n1 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_n1.log 2>&1
ncu -i $OUT/ncu_n1.ncu-rep --page raw --csv > $OUT/r2_ncu_traffic_n1.csv 2>/dev/null
ncu -i $OUT/ncu_n1.ncu-rep --page source --csv > $OUT/r2_ncu_source_n1.csv 2>/dev/null
tail -2 $OUT/ncu_n1.log
echo "== launch list n1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_launches_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
tail -3 $OUT/r2_launches_n1.csv
run_bench() { name=$1; shift
  echo "== bench $name"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "rc=$?"; head -c 600 $OUT/bench_$name.json; echo; grep -E "PARITY|Error|error|Traceback" $OUT/bench_$name.err | head -5
}
run_bench n2 --steps 30 --warmup 3
TSB_BENCH_PROFILE=1048576 TSB_SWEEP_BYTES=4294967296 run_bench cfg5 --config 5
TORCHSTORE_B200_FAST_LANE=0 TSB_SWEEP_BYTES=4294967296 run_bench cfg5_nofast --config 5
python - <<PY
import json
for f in ("cfg5","cfg5_nofast"):
    try:
        d=json.load(open("$OUT/bench_%s.json"%f))
        for r in d["config"]["sweep"]: print(f, r["key_bytes"]>>20, "MiB x", r["keys_per_batch"], "put %.0f get %.0f peer %s GB/s/GPU  put %.1f us/key" % (r["put_GBps_per_gpu"], r["get_local_GBps_per_gpu"], r["get_peer_GBps_per_gpu"], r["put_us_per_key"]))
    except Exception as e: print(f, e)
PY
grep -A 50 "cumulative" $OUT/bench_cfg5.err | head -70
ls -la $OUT
