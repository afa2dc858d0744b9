#!/bin/bash
cd "$(dirname "$0")/../.."
for lg in 16 20 22; do env ZKG_MSM_COOP_REDUCE=0 python tools/scratch/msm_reg.py $lg 5; env python tools/scratch/msm_reg.py $lg 5; done
env python tools/scratch/msm_reg.py 24 3; env ZKG_MSM_SORT_BPS_HIDDEN=1 python tools/scratch/msm_reg.py 24 3; env ZKG_MSM_SORT_BPS_HIDDEN=4 python tools/scratch/msm_reg.py 24 3
for lg in 16 19; do env ZKG_MSM_COOP_REDUCE=0 python tools/scratch/msm_reg_g2.py $lg 5; env python tools/scratch/msm_reg_g2.py $lg 5; done
