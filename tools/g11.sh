mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/pytest_multigpu.log 2>&1; echo "multigpu pytest rc=$?"; tail -3 gpurun_out/pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2>> gpurun_out/bench_2gpu.err; echo "ref2 rc=$?"; cat gpurun_out/bench_ref_2gpu.json
