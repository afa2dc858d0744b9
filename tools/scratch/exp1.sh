#!/bin/bash
# experiment 1: sort pipeline variants, block size, L2 fetch granularity (registered G1 2^22, 2^20; G2 2^19)
cd "$(dirname "$0")/../.."
run() { env "$@" python tools/scratch/msm_reg.py 22 5; }
echo "== baseline (no groups)"; run ZKG_MSM_GROUP0=0
for g in 2 3 4 6; do echo "== group0=$g side=1"; run ZKG_MSM_GROUP0=$g; done
echo "== group0=3 side=0"; run ZKG_MSM_GROUP0=3 ZKG_MSM_SIDE=0
echo "== TB64"; run ZKG_MSM_ACC_TB=64 ZKG_MSM_GROUP0=0
echo "== TB64 + group 3"; run ZKG_MSM_ACC_TB=64 ZKG_MSM_GROUP0=3
for f in 32 64 128; do echo "== L2 fetch $f"; run ZKG_L2_FETCH=$f ZKG_MSM_GROUP0=0; done
echo "== 2^20"; env ZKG_MSM_GROUP0=0 python tools/scratch/msm_reg.py 20 5; env python tools/scratch/msm_reg.py 20 5
echo "== 2^24"; env ZKG_MSM_GROUP0=0 python tools/scratch/msm_reg.py 24 3; env python tools/scratch/msm_reg.py 24 3
echo "== G2 2^19"; env ZKG_MSM_GROUP0=0 python tools/scratch/msm_reg_g2.py 19 5; env python tools/scratch/msm_reg_g2.py 19 5
