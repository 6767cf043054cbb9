#!/bin/bash
# A/B of encoder launch / ordering options: prints frames/s of the B=256 encode for each variant.
for v in "" "EMBCLIP_NO_PDL=1" "EMBCLIP_NO_SNAKE=1" "EMBCLIP_NO_PDL=1 EMBCLIP_NO_SNAKE=1"; do
  r=$(env $v python bench.py --steps 100 --warmup 10 --no-cpu --no-ppo --no-vit 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],3))")
  echo "[$v] $r"
done
