cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "conv or fused or adagn" > gpurun_out/p9_ops.log 2>&1; echo "ops rc=$?"; tail -n 4 gpurun_out/p9_ops.log
timeout 600 python -m pytest tests/test_gpu_network.py -q --no-header -p no:cacheprovider -x -m gpu -k "backbone_eps or sampler_trajectory or graph_replay" > gpurun_out/p9_net.log 2>&1; echo "net rc=$?"; tail -n 3 gpurun_out/p9_net.log
for cfg in "" "--fuse-adagn"; do
  name=$(echo "b256$cfg" | tr -d ' -')
  timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras $cfg > gpurun_out/p9_$name.json 2> gpurun_out/p9_$name.err; echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/p9_$name.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', 'conv frac', round(d['roofline']['frac'],3), {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])
except Exception as e: print('parse failed', e)
")"
done
IDF_OPTS="stats_item=0" timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras > gpurun_out/p9_win.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/p9_win.json').read().strip().splitlines()[-1]); print('window records:', round(d['value'],1),'img/s', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])"
