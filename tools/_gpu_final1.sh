mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/final/r2_gpu_tests.log 2>&1; tail -2 gpurun_out/final/r2_gpu_tests.log
python bench.py > gpurun_out/final/r2_bench.json 2> gpurun_out/final/r2_bench.err; tail -c 300 gpurun_out/final/r2_bench.json | head -c 300; echo
python bench.py --impl reference > gpurun_out/final/r2_bench_reference_arm.json 2> /dev/null; cut -c1-200 gpurun_out/final/r2_bench_reference_arm.json
python tools/size_sweep.py > gpurun_out/final/r2_size_sweep.jsonl 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/r2_launches.csv python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:plan_scan_philox_stream_kernel -c 2 -o gpurun_out/final/r2_scan python bench.py --paths 1e7 --steps 1 --warmup 0 --configs none --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:lsm_fused_step_kernel -s 100 -c 2 -o gpurun_out/final/r2_lsm python tools/ab_lsm.py ncu > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"gbm_slices_from_sums|pi_philox_stream|select_bracket_pass|gbm_terminal_affine" -c 12 -o gpurun_out/final/r2_misc python bench.py --paths 1e7 --steps 1 --warmup 0 --configs pi,portfolio_slices,stats --config-scale 0.1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/final
