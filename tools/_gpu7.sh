timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
for n in 1e6 3e6 1e7; do AB_PATHS=$n timeout 300 python tools/ab_lsm.py n$n 2>&1 | tail -1; done
python tools/size_sweep.py > gpurun_out/size_sweep_r2b.jsonl 2> gpurun_out/size_sweep.err; cat gpurun_out/size_sweep_r2b.jsonl
python tools/host_profile.py 1e6 8 2>&1 | grep -v INFO | head -40
ncu --set full --clock-control none --import-source on -k regex:lsm_fused_step_kernel -s 100 -c 3 -o gpurun_out/r2_lsm python tools/ab_lsm.py ncu > gpurun_out/ncu_lsm.log 2>&1; ls -la gpurun_out/r2_lsm.ncu-rep
