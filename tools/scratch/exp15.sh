#!/bin/bash
cd "$(dirname "$0")/../.."
for pf in 0 1 2; do env ZKG_MSM_PREFETCH=$pf ZKG_MSM_GROUP0=0 python tools/scratch/msm_reg.py 22 5; done
for pf in 0 1 2; do env ZKG_MSM_PREFETCH=$pf python tools/scratch/msm_reg.py 22 5; done
for pf in 0 1 2; do env ZKG_MSM_PREFETCH=$pf python tools/scratch/msm_reg_g2.py 19 5; done
for pf in 0 1; do env ZKG_MSM_PREFETCH=$pf python tools/scratch/msm_reg.py 20 5; done
