#!/bin/bash
# Round-2 GPU session F (2 GPUs): validation of the final tree -- full -m gpu suite, smoke, headline N=1/N=2, store sweep, reference arm.
set -u
OUT=gpurun_out/r2f
mkdir -p $OUT
echo "== pytest -m gpu (2 GPUs)"; timeout 1200 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu_2gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu_2gpu.log; tail -6 $OUT/pytest_gpu_2gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
run_bench() { name=$1; shift
  echo "== bench $name"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "rc=$?"; head -c 400 $OUT/bench_$name.json; echo; grep -E "PARITY|Error|error|Traceback" $OUT/bench_$name.err | head -5
}
run_bench n2 --steps 30 --warmup 3
TSB_SWEEP_BYTES=4294967296 run_bench cfg5 --config 5
echo "== bench n1"; timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; head -c 400 $OUT/bench_n1.json; echo
echo "== reference arm n1 (x2)"
for i in 1 2; do timeout 400 python bench.py --impl reference --steps 10 --warmup 1 > $OUT/bench_ref_n1_$i.json 2> $OUT/bench_ref_n1_$i.err; python -c "
import json; d=json.load(open('$OUT/bench_ref_n1_$i.json')); print('ref arm', d['value'], d['ms_per_step'], d['config']['ms_min'], d['config']['ms_max'], d['cpu_baseline']['cores'])"; done
python - <<PY
import json
d=json.load(open("$OUT/bench_cfg5.json"))
for r in d["config"]["sweep"]: print("cfg5", r["key_bytes"]>>20, "MiB x", r["keys_per_batch"], "put %.0f get %.0f peer %s GB/s/GPU" % (r["put_GBps_per_gpu"], r["get_local_GBps_per_gpu"], r["get_peer_GBps_per_gpu"]))
for f in ("n1","n2"):
    d=json.load(open("$OUT/bench_%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["e2e"]["value"], d.get("cpu_baseline") and d["cpu_baseline"]["value"])
PY
ls -la $OUT
