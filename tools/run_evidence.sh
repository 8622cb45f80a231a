#!/bin/bash
# Round-2 evidence: ncu --set full captures of every conv variant that carries time, the AdaGN kernel, a launch list of one
# UNet evaluation, the pipeline trace, the training-step profile.  Outputs under gpurun_out/ (copied to profiles/ by hand).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
R=${ROUND_TAG:-r2z}
cap() { local name=$1; shift; timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -o gpurun_out/${R}_ncu_$name -f python tools/prof_conv_one.py "$@" > gpurun_out/${R}_ncu_$name.log 2>&1; echo "ncu $name rc=$?"; }
cap conv_64x4_pair 64 64 64 0 0          # conv_halo_kernel<64,4,false,true>  (64->64 @64^2)
cap conv_128x2_pair 128 128 32 0 0       # conv_halo_kernel<128,2,false,true> (128->128 @32^2)
cap conv_128x1_pair 128 128 8 0 0        # conv_halo_kernel<128,*,false,true> (128->128 @8^2)
cap conv_64x4_pair_xf 64 64 64 0 1       # conv_halo_kernel<64,4,true,true>   (fused AdaGN)
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:adagn_apply -c 1 -o gpurun_out/${R}_ncu_adagn -f python tools/prof_adagn.py > gpurun_out/${R}_ncu_adagn.log 2>&1; echo "ncu adagn rc=$?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_unet_eval_b256.csv python tools/prof_step.py > gpurun_out/${R}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/list_launches.py > gpurun_out/${R}_launch_list_b256.txt 2>&1; echo "list rc=$?"
IDF_PROF_BATCH=32 timeout 300 python tools/list_launches.py > gpurun_out/${R}_launch_list_b32.txt 2>&1; echo "list32 rc=$?"
timeout 300 python tools/prof_train.py --graphs > gpurun_out/${R}_train_step_profile.txt 2>&1; echo "train prof rc=$?"
IDF_MB_XF=1 timeout 400 python tools/conv_microbench.py > gpurun_out/${R}_conv_microbench_b256.txt 2>&1; echo "mb rc=$?"
timeout 300 python tools/adagn_microbench.py > gpurun_out/${R}_adagn_microbench_b256.txt 2>&1; echo "adagn mb rc=$?"
if [ -f tools/ab/libidf_trace.so ]; then
  for c in "64 64 64 0 0" "64 64 64 1 0" "64 64 64 1 5" "64 64 64 1 1"; do echo "== conv_trace $c (cin cout H fused xf_debug)"; IDF_LIB_AB=tools/ab/libidf_trace.so timeout 100 python tools/conv_trace.py $c 2>&1 | tail -15; done > gpurun_out/${R}_conv_pipeline_trace.txt
fi
ls gpurun_out | grep ${R}_ | head -40
