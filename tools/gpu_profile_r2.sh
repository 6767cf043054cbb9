#!/bin/bash
# Round-2 evidence (1 GPU): full GPU tests, bench (all blocks), per-op table, ncu launch lists with DRAM bytes for the encoder step,
# the PPO update and one ViT forward, and `ncu --set full` captures of every tcgen05 kernel family + the cluster GRU kernels.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 900 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> /dev/null
timeout 300 python tools/profile_ops.py 256 $TAG > gpurun_out/ops_$TAG.txt 2>&1
timeout 300 python tools/profile_ac.py 128 60 > gpurun_out/ac_timing_$TAG.json 2>&1
timeout 300 python tools/small_batch_latency.py 8 60 > gpurun_out/small_batch_$TAG.json 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-ppo --no-vit"
N=$(python -c "
import sys; sys.path.insert(0, '.')
from embclip_b200.encoder import ClipRN50Encoder
from embclip_b200.synthetic import synthetic_rn50_state_dict
print(ClipRN50Encoder(synthetic_rn50_state_dict(), 'cuda:0').launches_per_forward(('trunk', 'avgpool', 'attnpool')))" 2>/dev/null | tail -1)
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 600 ncu --metrics $M --clock-control none -s $((3 * N)) -c $N --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
# PPO update: profile_ac.py runs pack(7) + forward(7+7) + loss(7) + backward(7) + update x4 ... : capture everything, summarise by kernel
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_ac_$TAG.csv python tools/profile_ac.py 128 60 > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_vit_$TAG.csv python tools/profile_vit.py 512 > /dev/null 2>&1
F="--set full --clock-control none --import-source on"
timeout 600 ncu $F -k regex:bneck_tail -s 15 -c 5 -o gpurun_out/prof_bneck_tail_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu $F -k regex:conv3x3_halo -s 39 -c 4 -o gpurun_out/prof_conv3x3_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu $F -k regex:stem_conv1_rows -s 3 -c 1 -o gpurun_out/prof_stem_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu $F -k regex:gemm2sm -s 40 -c 3 -o gpurun_out/prof_gemm2sm_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu $F -k regex:conv_gemm -s 60 -c 2 -o gpurun_out/prof_conv_gemm_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu $F -k regex:"attention_tc|gemm2sm" -s 30 -c 5 -o gpurun_out/prof_vit_$TAG -f python tools/profile_vit.py 512 > /dev/null 2>&1
timeout 600 ncu $F -k regex:"gru_cluster|wgrad_gemm" -s 6 -c 4 -o gpurun_out/prof_ac_$TAG -f python tools/profile_ac.py 128 60 > /dev/null 2>&1
# keep text summaries only: the .ncu-rep files exceed what gpurun copies back (64 MiB)
for f in gpurun_out/prof_*_$TAG.ncu-rep; do
  n=$(basename $f .ncu-rep)
  python tools/ncu_summary.py full $f gpurun_out/${n}_full.txt > /dev/null 2>&1
  ncu -i $f --page source --csv > gpurun_out/${n}_source.csv 2> /dev/null
  rm -f $f
done
python - <<PY
import glob, os
# source pages are large: keep the 60 hottest SASS/source lines per kernel by sampled stall count
for f in glob.glob("gpurun_out/prof_*_source.csv"):
    try:
        lines = open(f, errors="replace").read().splitlines()
        open(f, "w").write("\n".join(lines[:4000]))
    except Exception as e:
        print("trim failed", f, e)
PY
du -sh gpurun_out; ls -la gpurun_out | grep $TAG
