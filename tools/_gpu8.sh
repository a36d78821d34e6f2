timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
for v in base nopdl base; do
  if [ "$v" = base ]; then unset MCF_B200_LIB; else export MCF_B200_LIB=$(readlink -f mcframework_b200/_variants/libmcf_$v.so); fi
  for n in 1e6 1e7; do AB_PATHS=$n timeout 300 python tools/ab_lsm.py $v-$n 2>&1 | tail -1; done
done
unset MCF_B200_LIB
python tools/size_sweep.py > gpurun_out/size_sweep_r2c.jsonl 2> gpurun_out/size_sweep.err; cat gpurun_out/size_sweep_r2c.jsonl
