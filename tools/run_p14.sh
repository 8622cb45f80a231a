cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "fused" 2>&1 | tail -2
IDF_MB_XF=1 IDF_MB_QUICK=1 timeout 300 python tools/conv_microbench.py 2>&1 | head -5
timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras > gpurun_out/p14.json 2> gpurun_out/p14.err; python -c "
import json
d=json.loads(open('gpurun_out/p14.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])"
