#!/bin/bash
# Round-2 compute-sanitizer pass at HEAD (the r1 log predates the conv_gemm ring epilogue; racecheck had only ever run on the
# GRU): memcheck over every kernel family incl. the new ones (cluster GRU, torchvision plan, wide residual tile), racecheck on the
# hand-synchronised residual epilogue and the cluster GRU.
mkdir -p gpurun_out
run() {  # tool, tag, pytest args...
  local tool=$1 tag=$2; shift 2
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest "$@" -m gpu -x -q -p no:cacheprovider > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "[$tool $tag] exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitize_${tool}_${tag}.log | tr '\n' ' ')"
}
run memcheck prims tests/test_primitives_gpu.py -k "gemm_epilogues or gemm_cta_pair_epilogues or gemm_k_concat or gemm_residual_wide_tile or conv3x3 or stem_conv1 or avgpool2"
run memcheck tail tests/test_primitives_gpu.py -k "bneck_tail and not 148 and not 160"
run memcheck rn50 tests/test_rn50_gpu.py -k "golden or uint8"
run memcheck imagenet tests/test_imagenet_gpu.py
run memcheck act tests/test_actor_critic_gpu.py -k "act_step or encode_rows or harness_packed"
run memcheck ac tests/test_actor_critic_gpu.py -k "wgrad_transposed or (gru_forward_backward_vs_torch and not 128-60) or trainable or (forward_vs_oracle and not 128-60) or (ppo_loss_and_gradients_vs_oracle and not 128-60) or sumsq"
run memcheck storage tests/test_storage.py
run memcheck vit tests/test_vit_gpu.py -k "zero_shot or text_features"
run racecheck gru tests/test_actor_critic_gpu.py -k "(gru_forward_backward_vs_torch and (5-7-128 or 3-33-512)) or trainable"
run racecheck res tests/test_primitives_gpu.py -k "gemm_epilogues or gemm_residual_wide_tile"
run racecheck tailpool tests/test_primitives_gpu.py -k "bneck_tail_pool and (1-2-56 or 3-8-32)"
run racecheck stem tests/test_rn50_gpu.py -k "uint8"
run synccheck vit tests/test_vit_gpu.py -k "text_features"
