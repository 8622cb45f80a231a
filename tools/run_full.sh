cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=${ROUND_TAG:-r2n}
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${R}_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${R}_gputests.log
for cfg in "" "--fuse-adagn"; do
  name=$(echo "b256$cfg" | tr -d ' -')
  timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras $cfg > gpurun_out/${R}_$name.json 2> gpurun_out/${R}_$name.err; echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/${R}_$name.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', 'conv frac', round(d['roofline']['frac'],3), {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])
except Exception as e: print('parse failed', e)
")"
done
