#!/bin/bash
# ncu full capture (with source) of the full-size assembly + solve launches and of the many-path metrics kernel
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lm_assemble|lm_block_solve|path_metrics_many' -s 6 -c 3 -f -o gpurun_out/prof_r2_b python bench.py --chunks 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_b.log 2>&1
tail -2 gpurun_out/ncu_r2_b.log | cut -c1-200
