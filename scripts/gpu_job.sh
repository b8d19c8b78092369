# scratch driver for one gpurun call (edited per call; results land in gpurun_out/)
timeout 230 ncu --set full --clock-control none --import-source on -c 200 -o gpurun_out/ncu_smoke python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1; tail -2 gpurun_out/ncu_smoke.log | cut -c1-200; ls -la gpurun_out/ncu_smoke.ncu-rep | cut -c1-120
