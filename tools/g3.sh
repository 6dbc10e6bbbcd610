mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_umma_filter_gpu.py -m gpu -x -q -s -k "documented_bound" > gpurun_out/pytest_umma1.log 2>&1; echo "dump rc=$?"; tail -30 gpurun_out/pytest_umma1.log
timeout 600 python -m pytest tests/test_umma_filter_gpu.py -m gpu -x -q -k "not documented_bound" > gpurun_out/pytest_umma2.log 2>&1; echo "parity rc=$?"; tail -30 gpurun_out/pytest_umma2.log
timeout 300 python tools/tune_mma.py > gpurun_out/tune_mma.log 2>&1; grep -E "default|mma_cfg5|umma" gpurun_out/tune_mma.log
