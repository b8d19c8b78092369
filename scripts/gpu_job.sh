# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
bash scripts/gpu_checks.sh tests
