cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for o in "conv_l2_prefetch=0" "conv_l2_prefetch=1"; do echo "== $o"; IDF_OPTS="$o" IDF_MB_QUICK=1 timeout 300 python tools/conv_microbench.py 2>&1 | tail -10; done
