#!/bin/bash
cd "$(dirname "$0")/../.."
python tools/scratch/sweep_gen.py g1 8:8:8 10:8:8 12:10:10 13:11:15 14:15:15 16:15:15 18:15:15 20:15:16 22:17:17
python tools/scratch/sweep_gen.py g2 10:8:8 13:12:15 14:15:15 16:15:15 19:15:17
