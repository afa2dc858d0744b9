#!/bin/bash
cd "$(dirname "$0")/../.."
nproc
for rep in 1 2; do
for th in 7 3 11 15; do echo "threads=$th"; env ZKG_STAGING_THREADS=$th python tools/scratch/pageable_msm.py 22 2>&1 | grep "pageable"; done
done
