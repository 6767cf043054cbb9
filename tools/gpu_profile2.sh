#!/bin/bash
# Round-end evidence: tests, bench (all blocks), ncu launch lists + full captures of every tcgen05 kernel family.
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.json
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-ppo --no-vit"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 186 -c 62 --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 30 -c 2 -o gpurun_out/prof_conv3x3_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2sm -s 40 -c 3 -o gpurun_out/prof_gemm2sm_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 60 -c 2 -o gpurun_out/prof_conv_gemm_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 24 -c 1 -o gpurun_out/prof_attention_$TAG -f python tools/profile_vit.py 512 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_gemm|gru_" -s 10 -c 4 -o gpurun_out/prof_ac_$TAG -f python tools/profile_ac.py 128 60 > /dev/null 2>&1
ls -la gpurun_out | grep $TAG
