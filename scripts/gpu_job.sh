# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
echo "=== kernels+models"; timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -3 | cut -c1-600
echo "=== N=1 B=64 overlap on"; timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e 2>&1 | grep -E "decode:|prefill done|rror"
echo "=== N=1 B=64 overlap off"; MLXB200_OVERLAP_SHARED=0 timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e 2>&1 | grep -E "decode:|rror"
echo "=== N=1 B=256 overlap on"; timeout 300 python bench.py --steps 16 --warmup 3 --no-e2e --batch 256 2>&1 | grep -E "decode:|rror"
for B in 1 8; do echo "=== int4 B=$B splitk"; timeout 300 python bench.py --steps 32 --warmup 4 --no-e2e --batch $B --quant 4 2>&1 | grep -E "decode:|rror"; done
echo "=== bf16 B=1"; timeout 300 python bench.py --steps 32 --warmup 4 --no-e2e --batch 1 2>&1 | grep -E "decode:|rror"
echo "=== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -12 | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610"
echo "=== auto N=2"; timeout 600 $TR bench.py --gpus 2 --steps 16 --warmup 3 > gpurun_out/auto2.json 2> gpurun_out/auto2.log; grep -E "decode:|rror|Traceback" gpurun_out/auto2.log | sort | uniq | head -8 | cut -c1-200; grep "^{" gpurun_out/auto2.json | cut -c1-300
