#!/bin/bash
cd "$(dirname "$0")/../.."
for lg in 10 13 16 18; do env ZKG_MSM_REDUCE_L=8 python tools/scratch/msm_reg.py $lg 5; env python tools/scratch/msm_reg.py $lg 5; done
env python tools/scratch/msm_reg.py 22 5
for lg in 13 16; do env ZKG_MSM_REDUCE_L=8 python tools/scratch/msm_reg_g2.py $lg 5; env python tools/scratch/msm_reg_g2.py $lg 5; done
