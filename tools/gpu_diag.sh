#!/bin/bash
# First-contact GPU diagnostics: every test group in its own process (a trapped kernel poisons the CUDA
# context of the process it ran in), bounded by `timeout`, logs under gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/diag.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
for t in test_gemm_plain test_gemm_tile_variants test_gemm_epilogues test_gemm_k_concat test_gemm_grouped \
         test_conv3x3 test_avgpool2 test_stem_conv1 test_error_paths; do
  echo "=================== $t" >> $LOG
  timeout 300 python -m pytest tests/test_primitives_gpu.py -m gpu -q -k "$t" -p no:cacheprovider 2>&1 | tail -60 >> $LOG
done
for t in test_rn50_per_op_vs_fp16_path test_rn50_vs_fp32_oracle test_rn50_golden test_rn50_head_selection_and_determinism \
         test_rn50_rejects_bad_input test_rn50_large_batch_properties; do
  echo "=================== $t" >> $LOG
  timeout 600 python -m pytest tests/test_rn50_gpu.py -m gpu -q -s -k "$t" -p no:cacheprovider 2>&1 | tail -60 >> $LOG
done
grep -E "^(=====|[0-9]+ (passed|failed)|FAILED|ERROR)|passed|failed" $LOG | tail -60
