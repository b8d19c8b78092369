#!/bin/bash
# Round-2 evidence run (1 GPU, under gpurun): launch list of a decode step, ncu --set full of every decode kernel at full size,
# the FP8 expert GEMMs, every kernel of the smoke path, the split-K sweep, and the compute-sanitizer race / sync checks.
# Outputs land in gpurun_out/; the summaries derived from them are committed under profiles/.
#   gpurun --timeout 1500 -- 'bash scripts/profile_round2.sh'
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
STEP=${STEP:-all}
run() { [ "$STEP" = all ] || [ "$STEP" = "$1" ]; }

if run list; then   # per-launch device time of one decode step at the bench configuration (64 sequences, context 128)
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v4.csv \
      python scripts/launch_profile.py > gpurun_out/launches_v4.log 2>&1
fi
if run step; then   # every kernel of a decode step, full set, context 2048 (so the MLA decode kernel streams a real cache)
  timeout 600 $NCU --profile-from-start off -c 40 -o gpurun_out/ncu_decode_step python scripts/launch_profile.py --ctx 2048 > gpurun_out/ncu_decode_step.log 2>&1
  ncu -i gpurun_out/ncu_decode_step.ncu-rep --page raw --csv > gpurun_out/ncu_decode_step_raw.csv 2>/dev/null
  [ "$(stat -c %s gpurun_out/ncu_decode_step.ncu-rep 2>/dev/null || echo 0)" -gt 30000000 ] && rm -f gpurun_out/ncu_decode_step.ncu-rep
fi
if run fp8; then    # block-scaled FP8 expert GEMMs (kind::mxf8f6f4.block_scale)
  MLXB200_FP8_EXPERTS=1 timeout 400 $NCU --profile-from-start off -k regex:gemm_fp8 -c 2 -o gpurun_out/ncu_fp8 python scripts/launch_profile.py > gpurun_out/ncu_fp8.log 2>&1
fi
if run smoke; then  # coverage: every kernel launched by __graft_entry__.smoke() (prefill attention, sampler, KV writes, ...)
  timeout 600 $NCU -c 150 -o gpurun_out/ncu_smoke python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
  python scripts/ncu_summary.py gpurun_out/ncu_smoke.ncu-rep gpurun_out/ncu_smoke.md gpurun_out/ncu_smoke_raw.csv && rm -f gpurun_out/ncu_smoke.ncu-rep
fi
if run sweep; then
  timeout 300 tests/cuda/bin/gemm_sweep > gpurun_out/gemm_sweep.txt 2>&1
fi
if run sanitize; then
  K="linear or gated_up or moe_experts or moe_block or moe_route or rmsnorm or mla or fp8_gemm or sampler or flash_prefill"
  for tool in racecheck synccheck; do
    timeout 300 compute-sanitizer --tool $tool --error-exitcode 7 --log-file gpurun_out/$tool.log \
        python -m pytest tests/test_kernels_gpu.py tests/test_mla_gpu.py tests/test_fp8_gpu.py -m gpu -q -x -k "$K" > gpurun_out/${tool}_pytest.log 2>&1
    echo "$tool rc=$?" >> gpurun_out/sanitize_rc.txt
  done
fi
ls -la gpurun_out | tail -20
