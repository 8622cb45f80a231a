cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "conv or fused" > gpurun_out/p8_ops.log 2>&1; echo "ops rc=$?"; tail -n 2 gpurun_out/p8_ops.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline --no-extras --fuse-adagn > gpurun_out/p8_fuse.json 2> gpurun_out/p8_fuse.err; python -c "
import json
d=json.loads(open('gpurun_out/p8_fuse.json').read().strip().splitlines()[-1]); print(round(d['value'],1),'img/s', round(d['ms_per_step']/100,3),'ms/unet-step', 'conv frac', round(d['roofline']['frac'],3), {k: round(v['ms_per_unet_eval'],3) for k,v in (d.get('kernel_breakdown') or {}).items()}, d['clocks']['sm_mhz'])" || tail -3 gpurun_out/p8_fuse.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -o gpurun_out/p8_xf64 -f python tools/prof_conv_one.py 64 64 64 0 1 > gpurun_out/p8_ncu64.log 2>&1; echo "ncu64 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -o gpurun_out/p8_xf128 -f python tools/prof_conv_one.py 128 128 32 0 1 > gpurun_out/p8_ncu128.log 2>&1; echo "ncu128 rc=$?"
