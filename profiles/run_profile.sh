#!/bin/bash
# Usage (on the GPU box, via gpurun): profiles/run_profile.sh <tag> [bench args...]
# Writes gpurun_out/<tag>_launches.csv (per-launch device times) and gpurun_out/<tag>_count.ncu-rep (full set
# of the count kernel).  Numbers printed by bench.py under ncu are never bench values.
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 "$@" > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_count_ -s 3 -c 2 \
    -o gpurun_out/${tag}_count python bench.py --steps 2 --warmup 3 "$@" > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out/
