# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
echo "=== compute-sanitizer memcheck (kernel tests)"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 400 -k "linear or gated_up or moe_experts or moe_route or rmsnorm or mla_rope or paged_attention" 2>&1 | tail -3 | cut -c1-300
echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
BENCH="python bench.py --layers 3 --steps 1 --warmup 1 --no-graphs --no-e2e"
timeout 300 $NCU -k 'regex:gemm_persistent_kernel<\(int\)64, \(bool\)1' -s 2 -c 1 -o gpurun_out/ncu_experts_gateup_decode $BENCH > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log | cut -c1-150
timeout 300 $NCU -k 'regex:gemm_swapab_kernel<\(int\)64, \(bool\)0' -s 6 -c 1 -o gpurun_out/ncu_splitk_decode $BENCH > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log | cut -c1-150
timeout 300 $NCU -k 'regex:paged_attn_kernel' -s 4 -c 1 -o gpurun_out/ncu_attn_decode $BENCH > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log | cut -c1-150
echo "=== default bench N=1 (with e2e)"; timeout 400 python bench.py 2>gpurun_out/bench1.log | tee gpurun_out/bench1.json | cut -c1-1800
