cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for d in 0 1 2 3 4; do IDF_OPTS="adagn_impl=1" IDF_XF_DEBUG=$d python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --fuse-adagn > gpurun_out/q_r2f_xfdbg$d.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/q_r2f_xfdbg$d.json').read().strip().splitlines()[-1]); print('xf_debug $d', round(d['value'],1),'img/s', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()})"; done
