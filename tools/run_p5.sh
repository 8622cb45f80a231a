cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "conv or fused or adagn" > gpurun_out/p5_ops.log 2>&1; echo "ops rc=$?"; tail -n 3 gpurun_out/p5_ops.log
for d in 0 1 4; do IDF_XF_DEBUG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --fuse-adagn > gpurun_out/p5_xf$d.json 2>gpurun_out/p5_xf$d.err; python -c "
import json
d=json.loads(open('gpurun_out/p5_xf$d.json').read().strip().splitlines()[-1]); print('xf_debug $d', round(d['value'],1),'img/s', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])" || tail -3 gpurun_out/p5_xf$d.err; done
