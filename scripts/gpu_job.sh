# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
echo "=== EP tests"; timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 -x -k "expert_parallel" 2>&1 | tail -15 | cut -c1-800
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610"
echo "=== ep N=2"; timeout 300 $TR bench.py --gpus 2 --steps 16 --warmup 3 --parallelism ep --no-e2e 2>&1 | grep -E "decode:|experts |rror|Traceback" | sort | uniq | cut -c1-250 | head
