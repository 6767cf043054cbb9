#!/bin/bash
for v in "EMBCLIP_2SM=0" "EMBCLIP_2SM=1" "EMBCLIP_2SM=1 EMBCLIP_2SM_MINK=128" "EMBCLIP_2SM=1 EMBCLIP_2SM_MINK=512"; do
  r=$(env $v python bench.py --steps 100 --warmup 10 --no-cpu --no-ppo 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), 'vit', round(d['vit_zero_shot']['value']), round(d['vit_zero_shot']['ms_per_step'],2))")
  echo "[$v] $r"
done
