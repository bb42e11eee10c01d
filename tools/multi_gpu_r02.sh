#!/bin/bash
# Round-2 multi-GPU measurements (run through `gpurun --gpus N`): e2e A/B + the bench line at N GPUs.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tools/e2e_ab.py 2>/dev/null | tail -1 > gpurun_out/r02_e2e_ab_${N}gpu_zero_copy.json
D2D_HOST_COPY=1 $TR tools/e2e_ab.py 2>/dev/null | tail -1 > gpurun_out/r02_e2e_ab_${N}gpu_copy_engine.json
E2E_BIND=0 $TR tools/e2e_ab.py 2>/dev/null | tail -1 > gpurun_out/r02_e2e_ab_${N}gpu_zero_copy_unbound.json
$TR bench.py --gpus $N --steps 500 --warmup 20 2> gpurun_out/r02_bench_${N}gpu.err | tail -1 > gpurun_out/r02_bench_${N}gpu.json
cat gpurun_out/r02_e2e_ab_${N}gpu_*.json; cat gpurun_out/r02_bench_${N}gpu.json
