# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610"
echo "=== auto N=$N"; timeout 700 $TR bench.py --gpus $N --steps 24 --warmup 4 > gpurun_out/auto$N.json 2> gpurun_out/auto$N.log; grep -E "decode:|prefill done|rror|Traceback" gpurun_out/auto$N.log | sort | uniq | cut -c1-220 | head -12; grep "^{" gpurun_out/auto$N.json | cut -c1-3000
echo "=== ep_bench"; timeout 200 $TR bench/ep_bench.py --tokens 64 2>&1 | grep -E "^\{|rror" | tee gpurun_out/ep_bench$N.json | cut -c1-700
