cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=${ROUND_TAG:-r2n}
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${R}_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${R}_gputests.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${R}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${R}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "clocks", d["clocks"])
print({k: round(v["ms_per_unet_eval"],3) for k,v in d["kernel_breakdown"].items()})
for k in ("train","save_latent","ddpm1000","eager_gpu","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:600])
PY
