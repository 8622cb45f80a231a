cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=r2h
timeout 600 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "fused_adagn" > gpurun_out/q_${R}_ops.log 2>&1; echo "ops rc=$?"; tail -n 8 gpurun_out/q_${R}_ops.log
timeout 600 python -m pytest tests/test_gpu_network.py -q --no-header -p no:cacheprovider -x -m gpu -s -k "fused_adagn_mode" > gpurun_out/q_${R}_net.log 2>&1; echo "net rc=$?"; tail -n 8 gpurun_out/q_${R}_net.log
for o in "xf_ldg=1" "xf_ldg=0"; do
IDF_OPTS=$o timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --fuse-adagn > gpurun_out/q_${R}_b256.json 2> gpurun_out/q_${R}_b256.err; echo "$o rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_${R}_b256.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])
except Exception as e: print('parse failed', e)
")"; tail -3 gpurun_out/q_${R}_b256.err
done
IDF_OPTS="xf_ldg=1" timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --fuse-adagn --batch 32 > gpurun_out/q_${R}_b32.json 2> gpurun_out/q_${R}_b32.err; python -c "
import json
d=json.loads(open('gpurun_out/q_${R}_b32.json').read().strip().splitlines()[-1]); print('b32 fused ldg', round(d['value'],1),'img/s', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()})"
