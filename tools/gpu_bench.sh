#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench_${R}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${R}.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "conv TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "adagn GB/s", round(d["kernel_breakdown"]["adagn"]["achieved_GBps"],1), "clocks", d["clocks"], "cpu", d["cpu_baseline"])
print({k: round(v["ms_per_unet_eval"],3) for k,v in d["kernel_breakdown"].items()})
PY
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${R}.csv python tools/prof_step.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 4 -o gpurun_out/prof_conv_${R} -f python tools/prof_step.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:adagn -s 1 -c 3 -o gpurun_out/prof_adagn_${R} -f python tools/prof_step.py > gpurun_out/ncu_adagn.log 2>&1; echo "ncu adagn rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn -s 0 -c 1 -o gpurun_out/prof_attn_${R} -f python tools/prof_step.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
fi
ls -la gpurun_out | tail -n 8
