#!/bin/bash
# 2-GPU expert-parallel checks + A/B of the fused router/dispatch block (same box): gpurun --gpus 2 -- "bash scripts/ep_ab.sh"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 tests/mgpu/ep_stress.py fused 2>&1 | grep "rank 0\]\|EP_STRESS\|Error\|error" | head -20
timeout 150 $TR --master-port 29612 tests/mgpu/ep_model_parity.py 2>&1 | grep "mismatch\|EP_MODEL\|Error" | head
for v in 1 0 1 0; do MLXB200_EP_FUSED_ROUTE=$v timeout 300 $TR --master-port 2960$v bench.py --gpus 2 --steps 30 --warmup 5 --parallelism ep --no-baseline --no-e2e 2>&1 | grep "rank 0\] decode" | sed "s/^/[fused_route=$v] /"; done
