cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -x -m gpu -k "conv or downsample or upsample or tail or im2col" > gpurun_out/p1_ops.log 2>&1; echo "ops rc=$?"; tail -n 3 gpurun_out/p1_ops.log
timeout 400 python tools/conv_microbench.py > gpurun_out/p1_convmb.txt 2>&1; echo "mb rc=$?"; tail -12 gpurun_out/p1_convmb.txt
bash tools/run_p4.sh
