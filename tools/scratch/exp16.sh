#!/bin/bash
cd "$(dirname "$0")/../.."
for rep in 1 2; do
for plan in "" "4,16,32,48,64" "2,8,20,34,49,64" "8,24,44,64"; do echo "plan=[$plan]"; env ZKG_MSM_BOUNDS="$plan" python tools/scratch/pageable_msm.py 22 2>&1 | grep -v registered | head -3; done
done
