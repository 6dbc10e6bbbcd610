#!/bin/bash
# One gpurun call: GPU parity tests, the bench line (both arms), the ncu launch list of the bench
# command, one full ncu capture per hot kernel, the tuning tables and the pipe microbenchmarks.
# Everything lands in gpurun_out/; tools/make_profiles.py turns it into profiles/.  Development tool.
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench_${TAG}.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref_${TAG}.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 216 -c 300 --csv \
  --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 20 --warmup 3 --no-attack > gpurun_out/launches_bench.log 2>&1
for what in fwd bwd knn; do
  case $what in fwd) RX=nn_fwd;; bwd) RX=nn_bwd;; knn) RX=knn_kernel;; esac
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$RX -s 2 -c 1 -f \
    -o gpurun_out/${TAG}_${what} python tools/prof.py $what 50 > gpurun_out/ncu_${what}.log 2>&1
done
GA_TUNE=0=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_fwd -s 2 -c 1 -f \
  -o gpurun_out/${TAG}_fwdfp32 python tools/prof.py fwd 50 > gpurun_out/ncu_fwdfp32.log 2>&1
GA_TUNE=0=22 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_fwd -s 2 -c 1 -f \
  -o gpurun_out/${TAG}_fwdumma python tools/prof.py fwd 50 > gpurun_out/ncu_fwdumma.log 2>&1
timeout 120 python tools/umma_trace.py 50 > gpurun_out/umma_trace.txt 2>&1
timeout 200 python tools/tune_bwd.py > gpurun_out/tune_bwd.log 2>&1
timeout 300 python tools/tune_e2e.py > gpurun_out/tune_e2e.log 2>&1
timeout 300 python tools/tune_mma.py > gpurun_out/tune_mma.log 2>&1
timeout 120 tools/mmabench.bin > gpurun_out/mmabench.txt 2>&1
timeout 120 tools/tmembench.bin > gpurun_out/tmembench.txt 2>&1
cat gpurun_out/tmembench.txt gpurun_out/mmabench.txt
ls -la gpurun_out
