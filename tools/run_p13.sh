cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "adagn or fused" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_network.py -q --no-header -p no:cacheprovider -x -m gpu -k "backbone_eps or sampler_trajectory or graph_replay or fuse" 2>&1 | tail -2
for cfg in "" "--batch 32"; do
  name=$(echo "b$cfg" | tr -d ' -=')
  timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras $cfg > gpurun_out/p13_$name.json 2> gpurun_out/p13_$name.err; echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/p13_$name.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])
except Exception as e: print('parse failed', e)
")"
done
