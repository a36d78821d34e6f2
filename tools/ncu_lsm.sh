#!/bin/bash
# ncu captures of the Longstaff-Schwartz step kernel (under gpurun, one GPU; the program runs once WITHOUT ncu first):
# a single-pass DRAM-traffic list with the caches left alone (the L2 residency of the cash flows is part of the design),
# and one --set full capture.   bash tools/ncu_lsm.sh -> gpurun_out/r2_lsm_*.{csv,ncu-rep}
mkdir -p gpurun_out
python tools/ab_lsm.py plain > gpurun_out/r2_lsm_plain.log 2>&1 || exit 1
grep "^{" gpurun_out/r2_lsm_plain.log
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct"
timeout 600 ncu $M --clock-control none --cache-control none -k regex:lsm_fused_step_kernel -s 300 -c 60 --csv \
    --log-file gpurun_out/r2_lsm_traffic.csv python tools/ab_lsm.py ncu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:lsm_fused_step_kernel -s 300 -c 4 -f \
    -o gpurun_out/r2_lsm_full python tools/ab_lsm.py ncu > /dev/null 2>&1
ls -la gpurun_out/r2_lsm_full.ncu-rep
