#!/bin/bash
# quick kernel iteration: conv / AdaGN op tests, network parity, a short bench, the conv microbenchmark
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND_TAG:-r2b}
timeout 600 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "conv or adagn or downsample or upsample or tail or im2col" > gpurun_out/q_${R}_ops.log 2>&1; echo "ops rc=$?"; tail -n 5 gpurun_out/q_${R}_ops.log
timeout 600 python -m pytest tests/test_gpu_network.py -q --no-header -p no:cacheprovider -x -m gpu -k "backbone_eps or sampler_trajectory or graph_replay" > gpurun_out/q_${R}_net.log 2>&1; echo "net rc=$?"; tail -n 5 gpurun_out/q_${R}_net.log
for cfg in "" "--fuse-adagn" ${EXTRA_CFGS}; do
  name=$(echo "b256$cfg" | tr -d ' -')
  timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras $cfg > gpurun_out/q_${R}_$name.json 2> gpurun_out/q_${R}_$name.err; echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_${R}_$name.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', 'conv frac', round(d['roofline']['frac'],3), {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])
except Exception as e: print('parse failed', e)
")"
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --batch 32 > gpurun_out/q_${R}_b32.json 2> gpurun_out/q_${R}_b32.err; python -c "
import json
d=json.loads(open('gpurun_out/q_${R}_b32.json').read().strip().splitlines()[-1]); print('b32', round(d['value'],1),'img/s', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()})"
timeout 300 python tools/conv_microbench.py > gpurun_out/q_${R}_convmb.txt 2>&1; cat gpurun_out/q_${R}_convmb.txt | tail -12
