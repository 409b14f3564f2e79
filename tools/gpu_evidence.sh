#!/bin/bash
# round-2 evidence visit ($1 = tag): bench + reference arm, ncu launch list of the bench command, ncu full captures of
# every kernel.  gpurun brings back at most 64 MiB: the big reports are summarised on the box and deleted.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$1.json 2> gpurun_out/bench_ref_$1.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$1.log 2>&1
python profiles/summarize_ncu.py --launches gpurun_out/launches_$1.csv gpurun_out/$1_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lm_assemble|lm_block_solve' -s 6 -c 2 -f -o gpurun_out/prof_$1_lm python bench.py --chunks 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$1_lm.log 2>&1
python profiles/summarize_ncu.py gpurun_out/prof_$1_lm.ncu-rep gpurun_out/$1_fullsize_lm_kernels
timeout 1200 ncu --set full --clock-control none -k regex:'lm_assemble|lm_block_solve|lm_pose_step|collision_flags|path_metrics|dp_sweep|dp_mjac' -c 40 -f -o gpurun_out/prof_$1_all python tools/ncu_targets.py > gpurun_out/ncu_$1_all.log 2>&1
python profiles/summarize_ncu.py gpurun_out/prof_$1_all.ncu-rep gpurun_out/$1_all_kernels
rm -f gpurun_out/prof_$1_all.ncu-rep gpurun_out/launches_$1.csv
ls -la gpurun_out
