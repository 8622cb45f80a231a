cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=r2e
timeout 600 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "adagn" > gpurun_out/q_${R}_ops.log 2>&1; echo "ops rc=$?"; tail -n 5 gpurun_out/q_${R}_ops.log
for o in "adagn_impl=1" "adagn_impl=2" "adagn_impl=2,adagn_ctas2=592" "adagn_impl=2,adagn_ctas2=2368" "adagn_impl=2,adagn_ctas2=4736"; do echo "== $o"; IDF_OPTS=$o python tools/adagn_microbench.py 2>&1 | tail -3; done
for o in "adagn_impl=2" "adagn_impl=2,adagn_ctas2=592" "adagn_impl=2,adagn_ctas2=2368"; do
IDF_OPTS=$o timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras > gpurun_out/q_${R}_b256.json 2> gpurun_out/q_${R}_b256.err; echo "$o rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_${R}_b256.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', 'adagn frac', round(d['kernel_breakdown']['adagn']['frac'],3), {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])
except Exception as e: print('parse failed', e)
")"
done
IDF_OPTS="adagn_impl=2" timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --batch 32 > gpurun_out/q_${R}_b32.json 2> gpurun_out/q_${R}_b32.err; python -c "
import json
d=json.loads(open('gpurun_out/q_${R}_b32.json').read().strip().splitlines()[-1]); print('b32', round(d['value'],1),'img/s', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()})"
timeout 600 python -m pytest tests/test_gpu_network.py -q --no-header -p no:cacheprovider -x -m gpu > gpurun_out/q_${R}_net.log 2>&1; echo "net rc=$?"; tail -n 5 gpurun_out/q_${R}_net.log
