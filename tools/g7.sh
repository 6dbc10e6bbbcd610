mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_filter_gpu.py -m gpu -x -q > gpurun_out/pytest_umma.log 2>&1; echo "umma tests rc=$?"; tail -5 gpurun_out/pytest_umma.log
timeout 300 python tools/tune_mma.py > gpurun_out/tune_mma.log 2>&1; grep -E "mma_cfg5|umma" gpurun_out/tune_mma.log
timeout 120 python tools/umma_trace.py 50 > gpurun_out/umma_trace.txt 2>&1; head -30 gpurun_out/umma_trace.txt; tail -3 gpurun_out/umma_trace.txt
