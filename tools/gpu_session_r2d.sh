#!/bin/bash
# Round-2 GPU session D (8 GPUs): every BASELINE config through bench.py flags, bit-exact verify inside each.
set -u
OUT=gpurun_out/r2d
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
run_bench() { n=$1; name=$2; shift; shift
  echo "== bench $name (N=$n)"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "rc=$?"; head -c 700 $OUT/bench_$name.json; echo; grep -E "PARITY|Error|error|Traceback|deadline" $OUT/bench_$name.err | head -5
}
run_bench 8 n8 --steps 30 --warmup 3
run_bench 8 cfg4p --config 4p --steps 20 --warmup 3
run_bench 8 cfg3a --config 3a --steps 10 --warmup 3
run_bench 8 cfg3b --config 3b --steps 10 --warmup 3
run_bench 8 cfg3b_nccl --config 3b --allgather --steps 10 --warmup 3
run_bench 8 cfg5 --config 5
run_bench 4 n4 --steps 30 --warmup 3
python - <<PY
import json
for f in ("n8","cfg4p","cfg3a","cfg3b","cfg3b_nccl","n4"):
    try:
        d=json.load(open("$OUT/bench_%s.json"%f))
        print(f, "value %.0f GB/s  ms/step %.3f  kernel %.3f  roofline frac %.3f  e2e %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value")))
    except Exception as e: print(f, "ERR", e)
try:
    d=json.load(open("$OUT/bench_cfg5.json"))
    for r in d["config"]["sweep"]: print("cfg5", r["key_bytes"]>>20, "MiB x", r["keys_per_batch"], "put %.0f get %.0f peer %s GB/s/GPU" % (r["put_GBps_per_gpu"], r["get_local_GBps_per_gpu"], r["get_peer_GBps_per_gpu"]))
except Exception as e: print("cfg5 ERR", e)
PY
ls -la $OUT
