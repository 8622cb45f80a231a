cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for o in "conv_pair=1" "conv_pair=0"; do
for cfg in "--batch 32" "--batch 64"; do
  name=$(echo "b$cfg$o" | tr -d ' -=')
  IDF_OPTS="$o" timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras $cfg > gpurun_out/p12_$name.json 2> gpurun_out/p12_$name.err; echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/p12_$name.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()})
except Exception as e: print('parse failed', e)
")"
done; done
