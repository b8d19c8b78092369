# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610"
echo "=== auto N=$N"; timeout 500 $TR bench.py --gpus $N --steps 24 --warmup 4 > gpurun_out/auto$N.json 2> gpurun_out/auto$N.log; grep -E "decode:|rror|Traceback" gpurun_out/auto$N.log | sort | uniq | cut -c1-200 | head -8; grep "^{" gpurun_out/auto$N.json | cut -c1-400
echo "=== llama3-8b pp$N"; timeout 300 $TR bench.py --gpus $N --steps 24 --warmup 4 --model llama3-8b --no-e2e > gpurun_out/llama$N.json 2> gpurun_out/llama$N.log; grep -E "decode:|prefill done|rror|Traceback" gpurun_out/llama$N.log | sort | uniq | cut -c1-200 | head -6
