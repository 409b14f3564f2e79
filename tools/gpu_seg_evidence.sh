#!/bin/bash
# evidence visit for the segmented solve ($1 = tag): GPU suite, memcheck of the segmented kernels, bench, ncu full capture
# of the three passes at 1 and 1024 paths (summarised on the box), live per-pass times, pipeline step times.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_$1.log 2>&1; tail -3 gpurun_out/pytest_$1.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -q -x -k "segmented" > gpurun_out/memcheck_$1.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/memcheck_$1.log | tail -3
timeout 600 python bench.py --steps 20 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "bench rc=$?"
PATHS=1,1024 SEGS=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lm_seg' -s 3 -c 6 -f -o gpurun_out/prof_$1_seg python tools/ncu_seg_target.py > gpurun_out/ncu_$1_seg.log 2>&1
python profiles/summarize_ncu.py gpurun_out/prof_$1_seg.ncu-rep gpurun_out/$1_seg_kernels
rm -f gpurun_out/prof_$1_seg.ncu-rep
for n in 1 2 3; do CPPFLOW_SEG_PASSES=$n PATHS=1,16,256,1024,2048 SEGS=0,8,16,24 timeout 200 python tools/probe_segsolve.py > gpurun_out/seg_passes${n}_$1.jsonl 2>/dev/null; done
timeout 300 python tools/probe_segpipe.py > gpurun_out/segpipe_$1.jsonl 2>/dev/null
timeout 300 python tools/probe_plan.py > gpurun_out/plan_$1.log 2>&1
ls -la gpurun_out | tail -15
