#!/bin/bash
cd "$(dirname "$0")/../.."
for lg in 10 16 20 22; do env ZKG_MSM_COOP_TAIL=2 python tools/scratch/msm_gen.py $lg 5; env python tools/scratch/msm_gen.py $lg 5; done
for lg in 10 16 19; do env ZKG_MSM_COOP_TAIL=2 python tools/scratch/msm_gen.py $lg 5 g2; env python tools/scratch/msm_gen.py $lg 5 g2; done
