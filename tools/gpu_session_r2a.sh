#!/bin/bash
# Round-2 GPU session A (2 GPUs): parity of the new link path, full -m gpu suite incl. multi-GPU tests,
# 2-GPU emulation sweeps of the N=8/4/2 syncs, P2P microbench, ncu with NVLink counters, racecheck.
set -u
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
echo "== link smoke"; timeout 300 python -m pytest tests/test_gpu_copy_rects.py -x -q -k "link_ring or narrow_row or fenced or pool" > $OUT/link_smoke.log 2>&1; echo "rc=$?" >> $OUT/link_smoke.log; tail -3 $OUT/link_smoke.log
echo "== pytest -m gpu (2 GPUs)"; timeout 900 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu_2gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu_2gpu.log; tail -5 $OUT/pytest_gpu_2gpu.log
ENVS="TSB_LINK=0;TSB_LINK=1;TSB_LINK=1,TSB_LINK_STAGE_BYTES=4096,TSB_LINK_STAGES=8;TSB_LINK=1,TSB_LINK_STAGE_BYTES=16384,TSB_LINK_STAGES=4;TSB_LINK=1,TSB_LINK_STAGES=3;TSB_LINK=1,TSB_LINK_STAGES=8;TSB_LINK=1,TSB_CTAS_PER_SM=2;TSB_LINK=1,TSB_CTAS_PER_SM=4;TSB_LINK=0,TSB_CTAS_PER_SM=4"
echo "== sweep x2 n=8"; timeout 600 python tools/sweep_plan.py --mode x2 --n 8 --env "$ENVS" --out $OUT/sweep_x2_n8.json > $OUT/sweep_x2_n8.log 2>&1; tail -12 $OUT/sweep_x2_n8.log
echo "== sweep x2 n=4"; timeout 400 python tools/sweep_plan.py --mode x2 --n 4 --env "TSB_LINK=0;TSB_LINK=1;TSB_LINK=1,TSB_CTAS_PER_SM=2;TSB_LINK=1,TSB_LINK_STAGE_BYTES=16384,TSB_LINK_STAGES=4" --out $OUT/sweep_x2_n4.json > $OUT/sweep_x2_n4.log 2>&1; tail -5 $OUT/sweep_x2_n4.log
echo "== sweep x2 n=2"; timeout 400 python tools/sweep_plan.py --mode x2 --n 2 --env "TSB_LINK=0;TSB_LINK=1;TSB_LINK=1,TSB_CTAS_PER_SM=2;TSB_LINK=1,TSB_LINK_STAGE_BYTES=16384,TSB_LINK_STAGES=4" --out $OUT/sweep_x2_n2.json > $OUT/sweep_x2_n2.log 2>&1; tail -5 $OUT/sweep_x2_n2.log
echo "== p2p"; timeout 400 python tools/p2p_bench.py --out $OUT/p2p.json > $OUT/p2p.log 2>&1; tail -30 $OUT/p2p.log
echo "== ncu nvlink metrics"; ncu --query-metrics 2>/dev/null | grep -i -E "nvl|aperture_peer" | head -40 > $OUT/ncu_nvl_metrics.txt; wc -l $OUT/ncu_nvl_metrics.txt
for LINK in 1 0; do
  echo "== ncu x2 n=8 link=$LINK"
  TSB_LINK=$LINK timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 4 -c 2 -k regex:copy_rects \
      -o $OUT/ncu_x2_n8_link$LINK -f python tools/sweep_plan.py --mode x2 --n 8 --iters 1 > $OUT/ncu_x2_n8_link$LINK.log 2>&1
  tail -3 $OUT/ncu_x2_n8_link$LINK.log
done
echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_copy_rects.py -x -q -k "link_ring or narrow_row or many_tiny" > $OUT/racecheck.log 2>&1; echo "rc=$?" >> $OUT/racecheck.log; tail -8 $OUT/racecheck.log
echo "== memcheck"; timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_copy_rects.py -x -q -k "link_ring or narrow_row" > $OUT/memcheck.log 2>&1; echo "rc=$?" >> $OUT/memcheck.log; tail -5 $OUT/memcheck.log
ls -la $OUT
