cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for b in 32 64 128 256; do
IDF_PROF_BATCH=$b IDF_FUSE=1 CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/list_launches.py > gpurun_out/ll_r2j_b$b.txt 2>&1; echo "b=$b rc=$?"; tail -3 gpurun_out/ll_r2j_b$b.txt | cut -c1-300
done
