cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -q --no-header -p no:cacheprovider -m gpu -s -x -k "widths_diff_hard_wires" > gpurun_out/q_r2l_wide.log 2>&1; echo "wide rc=$?"; grep -E "passed|failed|FAILED|parity|Error|error" gpurun_out/q_r2l_wide.log | tail -12
timeout 900 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -m gpu -x > gpurun_out/q_r2l_ops.log 2>&1; echo "ops rc=$?"; tail -3 gpurun_out/q_r2l_ops.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras > gpurun_out/q_r2l_b256.json 2> gpurun_out/q_r2l_b256.err; python -c "
import json
d=json.loads(open('gpurun_out/q_r2l_b256.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', 'conv frac', round(d['roofline']['frac'],3), {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()})"
