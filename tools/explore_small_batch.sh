#!/bin/bash
# Round-2 exploration: where does the step time go at the strong-scaling shard sizes (32 / 64 / 128 images per GPU)?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND_TAG:-r2a}
run() { local name=$1; shift; timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras "$@" > gpurun_out/sb_${R}_$name.json 2> gpurun_out/sb_${R}_$name.err; echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/sb_${R}_$name.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()})
except Exception as e: print('parse failed', e)
")"; }
run b256
run b32 --batch 32
run b32_fuse --batch 32 --fuse-adagn
run b32_pdl --batch 32 --pdl 1
run b32_fuse_pdl --batch 32 --fuse-adagn --pdl 1
run b32_l2 --batch 32 --chunk 16 --lanes 2
run b32_l4 --batch 32 --chunk 8 --lanes 4
run b32_l2_fuse_pdl --batch 32 --chunk 16 --lanes 2 --fuse-adagn --pdl 1
run b64 --batch 64
run b64_fuse_pdl --batch 64 --fuse-adagn --pdl 1
run b128 --batch 128
IDF_PROF_BATCH=32 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${R}_b32.csv python tools/prof_step.py > gpurun_out/ncu_launches_b32.log 2>&1; echo "ncu launches rc=$?"
