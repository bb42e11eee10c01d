#!/bin/bash
# Probes the GPU box for the reference's third-party stack (SURVEY §8c / VERDICT item 1).
out=gpurun_out/probe_box.txt
{
echo "== imports"; for m in pymunk gym gymnasium stable_baselines3 pygame matplotlib Box2D; do python -c "import $m; print('$m', getattr($m,'__version__','?'), $m.__file__)" 2>&1 | tail -1; done
echo "== pip download"; timeout 60 python -m pip download --no-deps -d /tmp/pd pymunk 2>&1 | tail -2
echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "pymunk|gym|stable|chipmunk|pygame" ; ls /opt 2>/dev/null
echo "== find"; find / -xdev \( -iname "*pymunk*" -o -iname "*chipmunk*" -o -iname "cpSpaceStep*" -o -iname "stable_baselines3*" \) -not -path "/proc/*" 2>/dev/null | head -20
echo "== pip cache"; ls ~/.cache/pip 2>/dev/null | head; python -m pip config list 2>&1 | head
echo "== baseline/_ref"; ls baseline/_ref 2>/dev/null | head
echo "== cpu"; nproc; python -c "import os;print(len(os.sched_getaffinity(0)))"; lscpu | head -20
echo "== gpu"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
} > $out 2>&1
