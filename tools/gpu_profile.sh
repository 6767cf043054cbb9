#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of one bench step, one `--set full` capture of
# the dominant kernel.  Everything lands under gpurun_out/ (tag = $1); copy what should be judged to profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
python -m embclip_b200.build > gpurun_out/build_$TAG.log 2>&1 || python -c "import __graft_entry__ as g; g.build()" >> gpurun_out/build_$TAG.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
# launch list: 1 warm-up step skipped, then the launches of ~one step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-ppo --no-vit > gpurun_out/ncu_launch_$TAG.log 2>&1
# dominant kernel, full set, 3 launches from the middle of the network
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 60 -c 3 \
    -o gpurun_out/prof_conv_gemm_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu --no-ppo --no-vit > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -12
