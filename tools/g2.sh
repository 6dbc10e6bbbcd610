mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "host" > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_host.log
timeout 300 python tools/tune_e2e.py > gpurun_out/tune_e2e.log 2>&1; cat gpurun_out/tune_e2e.log
GA_TUNE=0=21 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_fwd -s 2 -c 1 -f -o gpurun_out/r01_fwdpersist python tools/prof.py fwd 50 > gpurun_out/ncu_fwdpersist.log 2>&1
timeout 200 python tools/tune_bwd.py > gpurun_out/tune_bwd.log 2>&1; cat gpurun_out/tune_bwd.log
