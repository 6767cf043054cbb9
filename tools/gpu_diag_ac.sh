#!/bin/bash
# Actor-critic path diagnostics: each test group in its own process (a trapped kernel poisons its CUDA context),
# bounded by `timeout`; log under gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/diag_ac.log
: > $LOG
for t in test_wgrad_vs_torch test_wgrad_transposed_store test_gru_forward_backward_vs_torch test_gru_rejects_bad_shapes \
         test_gae_vs_oracle test_adam_clip_vs_torch test_actor_critic_forward_vs_oracle test_ppo_loss_and_gradients_vs_oracle \
         test_autograd_surface_matches_fused_path test_ppo_update_vs_oracle test_full_size_block_properties; do
  echo "=================== $t" >> $LOG
  timeout 300 python -m pytest tests/test_actor_critic_gpu.py -m gpu -q -s -k "$t" -p no:cacheprovider 2>&1 | tail -${2:-45} >> $LOG
done
grep -E "^(=====|[0-9]+ (passed|failed)|FAILED|ERROR)|passed|failed" $LOG | tail -60
