#!/bin/bash
# Multi-GPU parity check + bench on N GPUs of one box (run under gpurun --gpus N):  bash tools/bench_multi_gpu.sh N
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | tail -3
SFX_E2E_DIAG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 10 --warmup 3 --cpu-baseline 0 2> gpurun_out/e2e_diag_$N.err | tail -1
grep "e2e rank" gpurun_out/e2e_diag_$N.err | sort | head -8
