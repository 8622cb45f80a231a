cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for o in "conv_l2_prefetch=1" "conv_l2_prefetch=0"; do
  IDF_OPTS="$o" timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --fuse-adagn > gpurun_out/p7_fuse.json 2> gpurun_out/p7_fuse.err; python -c "
import json
d=json.loads(open('gpurun_out/p7_fuse.json').read().strip().splitlines()[-1]); print('$o', round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', 'conv frac', round(d['roofline']['frac'],3), {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])" || tail -3 gpurun_out/p7_fuse.err
done
IDF_OPTS="conv_l2_prefetch=1" IDF_MB_XF=1 IDF_MB_QUICK=1 timeout 300 python tools/conv_microbench.py 2>&1 | tail -10
