# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29610"
echo "=== pp$N"; timeout 500 $TR bench.py --gpus $N --steps 24 --warmup 4 > gpurun_out/pp$N.json 2> gpurun_out/pp$N.log; grep -E "decode:|layers " gpurun_out/pp$N.log | sort | head -12 | cut -c1-200; cut -c1-300 gpurun_out/pp$N.json
echo "=== ep$N"; timeout 500 $TR bench.py --gpus $N --steps 24 --warmup 4 --parallelism ep > gpurun_out/ep$N.json 2> gpurun_out/ep$N.log; grep -E "decode:|prefill done|rror|Traceback" gpurun_out/ep$N.log | sort | head -6 | cut -c1-250; cut -c1-300 gpurun_out/ep$N.json
echo "=== ep$N B=256"; timeout 500 $TR bench.py --gpus $N --steps 24 --warmup 4 --parallelism ep --batch 256 --no-e2e > gpurun_out/ep${N}_b256.json 2> gpurun_out/ep${N}_b256.log; grep -E "decode:|prefill done|rror|Traceback" gpurun_out/ep${N}_b256.log | sort | head -4 | cut -c1-250
echo "=== ep_bench"; timeout 300 $TR bench/ep_bench.py --tokens 64 2>&1 | grep -E "^\{|rror" | tee gpurun_out/ep_bench$N.json | cut -c1-700
