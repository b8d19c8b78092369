#!/bin/bash
# Does the in-run baseline arm (which runs first) change the product's device-timed number?  Same box, three variants.
for rep in 1 2; do
  echo "[$rep] no baseline      :: $(timeout 300 python bench.py --no-baseline --no-e2e --steps 32 --warmup 4 2>&1 | grep 'rank 0\] decode')"
  echo "[$rep] baseline first   :: $(timeout 400 python bench.py --no-e2e --steps 32 --warmup 4 2>&1 | grep 'rank 0\] decode' | grep -v baseline)"
  echo "[$rep] baseline + 20 s  :: $(MLXB200_BENCH_SETTLE_S=20 timeout 400 python bench.py --no-e2e --steps 32 --warmup 4 2>&1 | grep 'rank 0\] decode' | grep -v baseline)"
done
