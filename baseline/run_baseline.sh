#!/bin/bash
# PyTorch + cuBLAS + NCCL-p2p re-implementation of the reference pipeline, 1/2/4/8 GPUs.
set -e
cd "$(dirname "$0")/.."
STEPS=${STEPS:-8}; WARMUP=${WARMUP:-3}
for N in ${GPUS:-1}; do
  if [ "$N" = "1" ]; then python bench.py --impl baseline --gpus 1 --steps $STEPS --warmup $WARMUP --no-e2e
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
         bench.py --impl baseline --gpus $N --steps $STEPS --warmup $WARMUP --no-e2e; fi
done
