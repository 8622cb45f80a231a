#!/bin/bash
# One GPU round: diagnostics + every GPU test group in its own process (a CUDA fault cannot poison the rest).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { local name=$1; shift; timeout 420 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "=== $name rc=$rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run diag python tools/diag_gpu.py
for t in test_layout_kernels test_conv3x3 test_conv3x3_forced_tiles_per_cta test_conv3x3_large_auto_tiles test_conv_epilogue_groupnorm_partials test_conv_data_gradient test_conv_weight_gradient test_adagn_backward test_conv1x1_qkv test_conv3x3_fused_shortcut test_downsample_stride2 test_upsample2x \
         test_im2col_head_and_gemm test_tail_fp32_and_sampler_epilogue test_adagn test_attention test_linear_and_gather \
         test_sampler_update test_mmd_against_oracle_and_golden; do
  TAILN=6 run ops_$t python -m pytest tests/test_gpu_ops.py -q --no-header -p no:cacheprovider -k "$t" -m gpu
done
for t in test_backbone_eps test_encoder test_sampler_trajectory test_reverse_ddim_both_variants \
         test_graph_replay_equals_eager_and_is_deterministic test_full_size_batch_independence_and_chunking test_loss_fn_forward_value test_training_step_gradients; do
  TAILN=14 run net_$t python -m pytest tests/test_gpu_network.py -q --no-header -p no:cacheprovider -s -k "$t" -m gpu
done
