# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
echo "=== multigpu tests"; timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -12 | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610"
echo "=== ep_bench breakdown T=64"; timeout 300 $TR bench/ep_bench.py --tokens 64 --breakdown 2>&1 | grep -E "^\{|rror|phases" | cut -c1-900
echo "=== auto N=2"; timeout 600 $TR bench.py --gpus 2 --steps 16 --warmup 3 > gpurun_out/auto2.json 2> gpurun_out/auto2.log; grep -E "decode:|rror|Traceback" gpurun_out/auto2.log | sort | uniq | head -8 | cut -c1-200; grep "^{" gpurun_out/auto2.json | cut -c1-300
echo "=== kernels"; timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -2 | cut -c1-400
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decode2.csv python bench.py --steps 1 --warmup 3 --no-graphs --no-e2e > gpurun_out/launch_bench.log 2>&1; wc -l gpurun_out/launches_decode2.csv
