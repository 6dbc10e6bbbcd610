mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "grad or bwd or backward or host" > gpurun_out/pytest_bwd.log 2>&1; echo "bwd tests rc=$?"; tail -4 gpurun_out/pytest_bwd.log
timeout 200 python tools/tune_bwd.py > gpurun_out/tune_bwd.log 2>&1; cat gpurun_out/tune_bwd.log
