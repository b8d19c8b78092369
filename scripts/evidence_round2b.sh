#!/bin/bash
# Second evidence run (1 GPU): coverage ncu of every repo kernel on the smoke path, racecheck on the shared-memory-heavy kernels,
# the full single-GPU test suite, smoke(), and the default bench.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 420 $NCU -k regex:b200 -c 90 -o gpurun_out/ncu_smoke python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_smoke.ncu-rep gpurun_out/ncu_smoke.md gpurun_out/ncu_smoke_raw.csv && rm -f gpurun_out/ncu_smoke.ncu-rep
K="gated_up or moe_block or mla_decode or fp8_gemm_dense or flash_prefill or sampler_greedy or rmsnorm"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/racecheck2.log \
    python -m pytest tests/test_kernels_gpu.py tests/test_mla_gpu.py tests/test_fp8_gpu.py -m gpu -q -x -k "$K" > gpurun_out/racecheck2_pytest.log 2>&1
echo "racecheck rc=$?" > gpurun_out/sanitize2_rc.txt
timeout 600 python -m pytest tests -m gpu -q --ignore=tests/test_multigpu.py 2>&1 | tail -5 | tee gpurun_out/gputests_1gpu_r2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2 | tee -a gpurun_out/gputests_1gpu_r2.log
timeout 500 python bench.py > gpurun_out/bench1_final.json 2> gpurun_out/bench1_final.log; tail -1 gpurun_out/bench1_final.json | cut -c1-600
timeout 200 python bench.py --impl reference | tail -1 | cut -c1-300
