#!/bin/bash
# Round-end evidence: tests, bench (all blocks), ncu launch lists + full captures of every tcgen05 kernel family.
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.json
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-ppo --no-vit"
# launch list of ONE step: skip the 3 warm-up steps, capture the launches of the timed one
N=$(python -c "
import sys; sys.path.insert(0, '.')
from embclip_b200.encoder import ClipRN50Encoder
from embclip_b200.synthetic import synthetic_rn50_state_dict
print(ClipRN50Encoder(synthetic_rn50_state_dict(), 'cuda:0').launches_per_forward(('trunk', 'avgpool', 'attnpool')))" 2>/dev/null | tail -1)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $((3 * N)) -c $N --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bneck_tail -s 3 -c 3 -o gpurun_out/prof_bneck_tail_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 30 -c 2 -o gpurun_out/prof_conv3x3_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2sm -s 40 -c 3 -o gpurun_out/prof_gemm2sm_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 60 -c 2 -o gpurun_out/prof_conv_gemm_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 24 -c 1 -o gpurun_out/prof_attention_$TAG -f python tools/profile_vit.py 512 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_gemm|gru_" -s 10 -c 4 -o gpurun_out/prof_ac_$TAG -f python tools/profile_ac.py 128 60 > /dev/null 2>&1
ls -la gpurun_out | grep $TAG
