mkdir -p gpurun_out
timeout 120 python tools/umma_trace.py 50 > gpurun_out/umma_trace.txt 2>&1; cat gpurun_out/umma_trace.txt
