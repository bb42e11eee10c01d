#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; each ncu pass follows a clean run of the same command).
set -x
export D2D_BENCH_MIN_S=0
CMD="python bench.py --no-extras --no-cpu-baseline --steps 40 --warmup 3"
$CMD > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:d2d_step_kernel -s 320 -c 3 -f -o gpurun_out/r02_step $CMD > gpurun_out/r02_ncu_step.log 2>&1
python tools/bench_agent.py 65536 large > gpurun_out/r02_agent_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:d2d_policy_kernel -s 400 -c 2 -f -o gpurun_out/r02_policy python tools/bench_agent.py 65536 large > gpurun_out/r02_ncu_policy.log 2>&1
ls -la gpurun_out/*.ncu-rep
