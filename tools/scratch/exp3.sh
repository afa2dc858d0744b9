#!/bin/bash
cd "$(dirname "$0")/../.."
run() { env "$@" python tools/scratch/msm_reg.py 22 5; }
echo "== no groups"; run ZKG_MSM_GROUP0=0
for g in 3 4; do for tb in 64 128; do for h in 1 2 4; do echo "== group0=$g tb=$tb hidden_bps=$h"; run ZKG_MSM_GROUP0=$g ZKG_MSM_SORT_BPS_HIDDEN=$h ZKG_MSM_SORT_TB_HIDDEN=$tb; done; done; done
