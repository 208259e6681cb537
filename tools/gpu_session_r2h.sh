#!/bin/bash
# Round-2 GPU session H (2 GPUs): latest tree -- contention evidence (solo vs concurrent), overhead after tight polling,
# context listing, init trace, store sweep with the single-reader peer get.
set -u
OUT=gpurun_out/r2h
mkdir -p $OUT
for N in 8 4 2; do
  timeout 300 python tools/sweep_plan.py --mode x2 --n $N --iters 5 --env "" --out $OUT/x2_n${N}_concurrent.json > $OUT/x2_n${N}_concurrent.log 2>&1
  timeout 300 python tools/sweep_plan.py --mode x2 --n $N --iters 5 --solo --env "" --out $OUT/x2_n${N}_solo.json > $OUT/x2_n${N}_solo.log 2>&1
  python - <<PY
import json
a=json.load(open("$OUT/x2_n${N}_concurrent.json"))[0]; b=json.load(open("$OUT/x2_n${N}_solo.json"))[0]
print("N=$N concurrent %.4f ms (nvlink %.0f GB/s)   peer idle %.4f ms (nvlink %.0f GB/s)" % (a["ms_median"], a["nvlink_in_GBps"], b["ms_median"], b["nvlink_in_GBps"]))
PY
done
run_bench() { name=$1; shift
  echo "== bench $name"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "rc=$?"; grep -E "PARITY|Error|error|Traceback" $OUT/bench_$name.err | head -5
}
TSB_BENCH_LIST_CONTEXTS=1 TSB_TRACE_INIT=1 run_bench n2 --steps 30 --warmup 3
TSB_SWEEP_BYTES=4294967296 run_bench cfg5 --config 5
echo "== bench n1"; timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"
python - <<PY
import json
for f in ("n1","n2"):
    d=json.load(open("$OUT/bench_%s.json"%f)); print(f, "value %.0f ms/step %.4f kernel %.4f overhead %.4f e2e %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["config"]["host_overhead_ms_per_step"], d["e2e"]["value"]))
d=json.load(open("$OUT/bench_cfg5.json"))
for r in d["config"]["sweep"]: print("cfg5", r["key_bytes"]>>20, "MiB x", r["keys_per_batch"], "put %.0f get %.0f peer(all) %.0f peer(solo) %.0f GB/s/GPU" % (r["put_GBps_per_gpu"], r["get_local_GBps_per_gpu"], r["get_peer_GBps_per_gpu"], r["get_peer_single_reader_GBps"]))
PY
grep -E "spmd r0|state dicts wrapped|store initialized|weights ready" $OUT/bench_n2.err | head
grep -A 6 "compute contexts" $OUT/bench_n2.err | head -10; nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader
