# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | cut -c1-300
for Q in 0 4; do for B in 64 256; do echo "=== quant=$Q B=$B"; timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e --batch $B --quant $Q 2>&1 | grep -E "decode:|rror"; done; done
echo "=== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -15 | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610"
echo "=== boundary"; timeout 300 $TR bench/boundary_bench.py 2>&1 | grep -E "^\{|rror" | cut -c1-1500
echo "=== ep_bench T=64"; timeout 300 $TR bench/ep_bench.py --tokens 64 2>&1 | grep -E "^\{|rror" | cut -c1-1200
echo "=== ep_bench T=256"; timeout 300 $TR bench/ep_bench.py --tokens 256 2>&1 | grep -E "^\{|rror" | cut -c1-1200
echo "=== pp2"; timeout 400 $TR bench.py --gpus 2 --steps 16 --warmup 3 2>&1 | grep -E "decode:|^\{|rror" | cut -c1-1800
echo "=== ep2"; timeout 400 $TR bench.py --gpus 2 --steps 16 --warmup 3 --parallelism ep 2>&1 | grep -E "decode:|prefill done|^\{|rror|Traceback" | cut -c1-1800
