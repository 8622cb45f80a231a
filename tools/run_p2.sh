cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -o gpurun_out/p2_conv64 -f python tools/prof_conv_one.py 64 64 64 > gpurun_out/p2_ncu64.log 2>&1; echo "ncu64 rc=$?"; tail -3 gpurun_out/p2_ncu64.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -o gpurun_out/p2_conv128 -f python tools/prof_conv_one.py 128 128 32 > gpurun_out/p2_ncu128.log 2>&1; echo "ncu128 rc=$?"; tail -3 gpurun_out/p2_ncu128.log
