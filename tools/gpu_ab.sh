#!/bin/bash
# A/B visit: parity tests, kernel timings, bench at the driver's step count ($1 = tag of the output files)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_$1.log; cat gpurun_out/pytest_$1.log
timeout 300 python tools/probe_kernels.py > gpurun_out/probe_$1.log 2>&1; cat gpurun_out/probe_$1.log
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['ms_per_step','ms_per_step_without_tail','single_stream_ms_per_step']}, d['kernel_ms']['lm_assemble_kernel'], d['kernel_ms']['lm_block_solve_kernel'], d['e2e']['ms_per_step'], d['strong_preview']['ms_per_step'], d['roofline_step']['frac'])
PY
