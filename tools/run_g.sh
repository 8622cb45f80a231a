cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2g.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2g.json").read().strip().splitlines()[-1])
for k in ("value","scaling","e2e","eager_gpu","cpu_baseline","save_latent","ddpm1000","train"):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 900 python -m pytest tests/test_gpu_variants.py -q --no-header -p no:cacheprovider -x -m gpu -k "run_py" > gpurun_out/q_r2g_runpy.log 2>&1; echo "runpy rc=$?"; tail -n 15 gpurun_out/q_r2g_runpy.log
