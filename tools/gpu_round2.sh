#!/bin/bash
# Round-2 evidence in one gpurun call: GPU parity tests, smoke, both bench arms, the ncu launch list of the bench
# command and one full ncu capture per hot kernel.  Outputs in gpurun_out/; tools/make_profiles.py summarises them.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench_${TAG}.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref_${TAG}.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 216 -c 300 --csv \
  --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 20 --warmup 3 --no-attack --no-legs > gpurun_out/launches_bench.log 2>&1
for what in fwd bwd knn; do
  case $what in fwd) RX=nn_fwd;; bwd) RX=nn_bwd;; knn) RX=knn_;; esac
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$RX -s 2 -c 1 -f \
    -o gpurun_out/${TAG}_${what} python tools/prof.py $what 50 > gpurun_out/ncu_${what}.log 2>&1
done
ls -la gpurun_out | tail -20
