#!/bin/bash
# Per-GPU-batch sweep of the sampler's micro-batch lanes (development helper): does spreading a small shard over two
# streams lift the launch-latency floor that bounds strong scaling?
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for cfg in "32 0 1" "32 16 2" "32 8 4" "64 0 1" "64 32 2" "64 16 4" "128 64 2" "256 128 2"; do
  set -- $cfg
  extra=""; [ "$2" != "0" ] && extra="--chunk $2 --lanes $3"
  timeout 300 python bench.py --batch $1 $extra --steps 3 --warmup 3 --no-extras --no-train --no-cpu-baseline 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('batch $1 chunk $2 lanes $3 ->', round(d['value'],1), 'img/s  e2e', round(d['e2e']['value'],1), 'launches', d.get('gpu_launches'))"
done
