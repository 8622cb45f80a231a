cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for h in 64 32; do
  IDF_FUSE_MIN_H=$h timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras > gpurun_out/q1_$h.json 2> gpurun_out/q1_$h.err; echo "min_h $h rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/q1_$h.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])
except Exception as e: print('parse failed', e)
")"
done
