#!/bin/bash
# quick GPU visit: parity tests, timing probe, optional ncu capture ($1 = kernel regex, empty = skip)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_kernels.py > gpurun_out/probe.log 2>&1; cat gpurun_out/probe.log
if [ -n "$1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 6 -c 2 -f -o gpurun_out/prof_q python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
fi
