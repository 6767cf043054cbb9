TAG=r2j
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-ppo --no-vit"
N=$(python -c "
import sys; sys.path.insert(0, '.')
from embclip_b200.encoder import ClipRN50Encoder
from embclip_b200.synthetic import synthetic_rn50_state_dict
print(ClipRN50Encoder(synthetic_rn50_state_dict(), 'cuda:0').launches_per_forward(('trunk', 'avgpool', 'attnpool')))" 2>/dev/null | tail -1)
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 600 ncu --metrics $M --clock-control none -s $((3 * N)) -c $N --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_ac_$TAG.csv python tools/profile_ac.py 128 60 > /dev/null 2>&1
F="--set full --clock-control none --import-source on"
timeout 600 ncu $F -k regex:bneck_tail -s 15 -c 5 -o gpurun_out/prof_bneck_tail_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu $F -k regex:conv3x3_halo -s 39 -c 4 -o gpurun_out/prof_conv3x3_$TAG -f $B > /dev/null 2>&1
timeout 600 ncu $F -k regex:attnpool_tc -s 3 -c 1 -o gpurun_out/prof_attnpool_$TAG -f $B > /dev/null 2>&1
for f in gpurun_out/prof_*_$TAG.ncu-rep; do
  n=$(basename $f .ncu-rep)
  python tools/ncu_summary.py full $f gpurun_out/${n}_full.txt > /dev/null 2>&1
  ncu -i $f --page source --csv 2>/dev/null | head -4000 > gpurun_out/${n}_source.csv
  rm -f $f
done
timeout 300 python tools/profile_ac.py 128 60 > gpurun_out/ac_timing_$TAG.json 2>&1
timeout 300 python tools/small_batch_latency.py 8 60 > gpurun_out/small_batch_$TAG.json 2>&1
ls -la gpurun_out | grep $TAG
