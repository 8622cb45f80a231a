#!/bin/bash
# driver-style 2-GPU launch of the bench (strong scaling headline, weak scaling, data-parallel training, dp_check)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=${ROUND_TAG:-r2z}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err; echo "rc=$?"; tail -3 gpurun_out/${R}_bench_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_bench_n2.json").read().strip().splitlines()[-1])
for k in ("value","scaling","config","e2e","weak_scaling","train","dp_check"):
    print(k, json.dumps(d.get(k))[:500])
PY
