#!/bin/bash
# ncu captures of the round-2 statistics kernels (under gpurun, one GPU; each program runs once WITHOUT ncu first).
#   bash tools/ncu_stats_kernels.sh   -> gpurun_out/r2_boot_*.ncu-rep, r2_select.ncu-rep, launch lists
set -u
mkdir -p gpurun_out
python tools/prof_boot.py 1e8 64 single > gpurun_out/r2_boot_plain.log 2>&1 || exit 1
python tools/prof_select.py > gpurun_out/r2_select_plain.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:'boot_bin_safe_kernel|boot_fuse_verify_kernel|boot_count_kernel|boot_bin_kernel' -c 8 -f -o gpurun_out/r2_boot_bin \
    python tools/prof_boot.py 1e8 64 single > gpurun_out/r2_boot_bin.log 2>&1
timeout 900 $NCU -k regex:boot_gather_kernel -s 24 -c 3 -f -o gpurun_out/r2_boot_gather python tools/prof_boot.py 1e8 64 single > gpurun_out/r2_boot_gather.log 2>&1
timeout 900 $NCU -k regex:'select_sample_kernel|select_bracket' -s 8 -c 8 -f -o gpurun_out/r2_select python tools/prof_select.py > gpurun_out/r2_select.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_boot_launches.csv \
    python tools/prof_boot.py 1e8 64 single > gpurun_out/r2_boot_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_select_launches.csv \
    python tools/prof_select.py 2>&1 > gpurun_out/r2_select_launches.log
cat gpurun_out/r2_boot_plain.log gpurun_out/r2_select_plain.log | tail -6
ls -la gpurun_out/*.ncu-rep
