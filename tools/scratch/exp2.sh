#!/bin/bash
cd "$(dirname "$0")/../.."
run() { env "$@" python tools/scratch/msm_reg.py 22 5; }
echo "== no groups, BPS 32 / 8 / 0(uncapped)"; run ZKG_MSM_GROUP0=0; run ZKG_MSM_GROUP0=0 ZKG_MSM_SORT_BPS=8; run ZKG_MSM_GROUP0=0 ZKG_MSM_SORT_BPS=0
for g in 2 3 4; do for h in 1 2 4; do echo "== group0=$g hidden_bps=$h"; run ZKG_MSM_GROUP0=$g ZKG_MSM_SORT_BPS_HIDDEN=$h; done; done
echo "== 2^20"; env python tools/scratch/msm_reg.py 20 5
echo "== 2^24"; env python tools/scratch/msm_reg.py 24 3
echo "== G2 2^19"; env python tools/scratch/msm_reg_g2.py 19 5
