"""Grid of launch settings for the 2-GPU emulation of the N-rank sync (tools/sweep_plan.py --mode x2):
link queue on/off x CTAs per SM x loads in flight per thread x ring geometry.  Prints the env string
for --env.  (TSB_COPY_UNROLL was a knob of the round-2 experiment builds -- 2 loads in flight per copy
thread measured no gain and was removed; the shipped library ignores it.)"""
import itertools
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
envs = []
for ctas, unroll in itertools.product((1, 2, 3), (2, 4)):
    envs.append(f"TSB_LINK=0,TSB_CTAS_PER_SM={ctas},TSB_COPY_UNROLL={unroll}")
for ctas, unroll, stages, sb in itertools.product((1, 2, 3), (2, 4), (3, 4, 6), (4096, 8192)):
    if n != 8 and (stages == 4 or (ctas == 1 and unroll == 2)):
        continue
    envs.append(f"TSB_LINK=1,TSB_CTAS_PER_SM={ctas},TSB_COPY_UNROLL={unroll},TSB_LINK_STAGES={stages},TSB_LINK_STAGE_BYTES={sb}")
print(";".join(envs))
