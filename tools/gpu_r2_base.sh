#!/bin/bash
# round-2 baseline visit: parity tests, bench at the driver's step count, fresh full capture of the full-size LM kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/bench_r2c.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lm_assemble|lm_block_solve' -s 6 -c 2 -f -o gpurun_out/prof_r2_full python bench.py --chunks 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_full.log 2>&1
tail -2 gpurun_out/ncu_r2_full.log | cut -c1-200
