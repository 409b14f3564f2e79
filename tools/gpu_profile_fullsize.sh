#!/bin/bash
# ncu --set full captures of the FULL-SIZE kernels (P = 8192, single stream): the chunk-pipelined bench launches 2048-path
# chunks, whose per-launch times say little about the roofline kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lm_assemble|lm_block_solve' -s 6 -c 2 -f -o gpurun_out/prof_full_ring4 python bench.py --chunks 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_ring4.log 2>&1
tail -2 gpurun_out/ncu_full_ring4.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lm_block_solve|path_metrics|lm_pose_step|dp_sweep' -s 4 -c 6 -f -o gpurun_out/prof_full_ring3 python tools/probe_kernels.py > gpurun_out/ncu_full_ring3.log 2>&1
tail -2 gpurun_out/ncu_full_ring3.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
