#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; echo "bench rc=$?"; cat gpurun_out/bench_${R}.json; tail -n 5 gpurun_out/bench_${R}.err
for c in 128 64 32; do
  python bench.py --steps 2 --warmup 3 --chunk $c --no-cpu-baseline > gpurun_out/bench_${R}_chunk$c.json 2> gpurun_out/bench_${R}_chunk$c.err; echo "bench chunk $c rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${R}_chunk$c.json"))
print("chunk $c value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "conv TF/s", round(d["roofline"]["achieved"],1), "adagn GB/s", round(d["kernel_breakdown"]["adagn"]["achieved_GBps"],1))
print({k: round(v["ms_per_unet_eval"],3) for k,v in d["kernel_breakdown"].items()})
PY
done
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${R}.csv python tools/prof_step.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm -s 12 -c 3 -o gpurun_out/prof_conv_${R} -f python tools/prof_step.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:adagn -s 6 -c 2 -o gpurun_out/prof_adagn_${R} -f python tools/prof_step.py > gpurun_out/ncu_adagn.log 2>&1; echo "ncu adagn rc=$?"
ls -la gpurun_out | tail -n 12
