export PYTHONDONTWRITEBYTECODE=1
ncu --set full --clock-control none --import-source on -k regex:'dwconv_bwd' -s 4 -c 1 -f -o gpurun_out/r01e_dwbwd_s0 python tools/kbench.py --only dwconv_bwd --stages 0 --reps 3 > gpurun_out/ncu_k0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dwconv_bwd' -s 4 -c 1 -f -o gpurun_out/r01e_dwbwd_s2 python tools/kbench.py --only dwconv_bwd --stages 2 --reps 3 > gpurun_out/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dwconv_fwd' -s 4 -c 1 -f -o gpurun_out/r01e_dwfwd_s0 python tools/kbench.py --only dwconv_fwd --stages 0 --reps 3 > gpurun_out/ncu_k3.log 2>&1
ls -la gpurun_out/*.ncu-rep
