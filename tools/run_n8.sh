#!/bin/bash
# driver-style 8-GPU launch of the bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=${ROUND_TAG:-r2z}; N=${NGPU:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err; echo "rc=$?"; tail -3 gpurun_out/${R}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_bench_n$N.json").read().strip().splitlines()[-1])
for k in ("value","scaling","e2e","weak_scaling","train","dp_check"):
    print(k, json.dumps(d.get(k))[:400])
PY
