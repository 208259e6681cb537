#!/bin/bash
# Round-2 GPU session J (2 GPUs): final validation + the staging experiment on the REAL symmetric N=2 job.
set -u
OUT=gpurun_out/r2j
mkdir -p $OUT
echo "== pytest -m gpu (2 GPUs)"; timeout 1200 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu_2gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu_2gpu.log; tail -4 $OUT/pytest_gpu_2gpu.log
run_bench() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name rc=$?"; grep -E "PARITY|Error|error|Traceback" $OUT/bench_$name.err | head -3
}
run_bench n2 --steps 30 --warmup 3
TORCHSTORE_B200_STAGE=1 TSB_L2_STREAM=1 TSB_STAGE_LEAD_BYTES=4194304 run_bench n2_stage4m --steps 30 --warmup 3 --no-e2e
TORCHSTORE_B200_STAGE=1 TSB_L2_STREAM=1 TSB_STAGE_LEAD_BYTES=8388608 run_bench n2_stage8m --steps 30 --warmup 3 --no-e2e
TORCHSTORE_B200_STAGE=1 TSB_STAGE_LEAD_BYTES=2097152 run_bench n2_stage2m_nostream --steps 30 --warmup 3 --no-e2e
echo "== bench n1"; timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"
echo "== reference arm"; timeout 400 python bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err; echo "rc=$?"
python - <<PY
import json
for f in ("n1","n2","n2_stage4m","n2_stage8m","n2_stage2m_nostream"):
    try:
        d=json.load(open("$OUT/bench_%s.json"%f)); print(f, "value %.0f ms/step %.4f kernel %.4f overhead %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["config"]["host_overhead_ms_per_step"]))
    except Exception as e: print(f, "ERR", e)
d=json.load(open("$OUT/bench_ref_n2.json")); print("ref n2", d["value"], d["cpu_baseline"]["gloo_transport"])
PY
