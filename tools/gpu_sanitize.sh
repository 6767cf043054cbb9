#!/bin/bash
# compute-sanitizer over small-shape parity tests of every kernel family (memcheck, then racecheck / synccheck on the
# hand-synchronised kernels).  Logs under gpurun_out/sanitize_*.log; summary lines at the end.
mkdir -p gpurun_out
T="tests/test_primitives_gpu.py -k gemm_plain_or_gemm_cta_pair_epilogues_or_conv3x3_or_stem"
run() {  # tool, tag, pytest args...
  local tool=$1 tag=$2; shift 2
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest "$@" -m gpu -x -q -p no:cacheprovider > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "[$tool $tag] exit $? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tool}_${tag}.log | tr '\n' ' ')"
}
run memcheck prims tests/test_primitives_gpu.py -k "gemm_epilogues or gemm_cta_pair_epilogues or gemm_k_concat or conv3x3 or stem_conv1 or avgpool2"
run memcheck tail tests/test_primitives_gpu.py -k "bneck_tail and not 148"
run memcheck rn50 tests/test_rn50_gpu.py -k "golden or uint8"
run memcheck act tests/test_actor_critic_gpu.py -k "act_step or encode_rows or harness_packed"
run memcheck ac tests/test_actor_critic_gpu.py -k "wgrad_transposed or gru_forward_backward_vs_torch or forward_vs_oracle or ppo_loss_and_gradients_vs_oracle"
run memcheck vit tests/test_vit_gpu.py -k "zero_shot or text_features"
run racecheck ac tests/test_actor_critic_gpu.py -k "gru_forward_backward_vs_torch and 5-7-128"
run synccheck vit tests/test_vit_gpu.py -k "text_features"
