mkdir -p gpurun_out
GA_TUNE=0=22 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nn_fwd -s 2 -c 1 -f -o gpurun_out/r01_fwdumma python tools/prof.py fwd 50 > gpurun_out/ncu_fwdumma.log 2>&1; tail -3 gpurun_out/ncu_fwdumma.log
