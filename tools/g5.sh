mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_filter_gpu.py -m gpu -x -q > gpurun_out/pytest_umma.log 2>&1; echo "umma tests rc=$?"; tail -5 gpurun_out/pytest_umma.log
timeout 300 python tools/tune_mma.py > gpurun_out/tune_mma.log 2>&1; grep -E "mma_cfg5|umma" gpurun_out/tune_mma.log
GA_TUNE=0=22 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_fwd -s 2 -c 1 -f -o gpurun_out/r01_fwdumma python tools/prof.py fwd 50 > gpurun_out/ncu_fwdumma.log 2>&1; tail -2 gpurun_out/ncu_fwdumma.log
