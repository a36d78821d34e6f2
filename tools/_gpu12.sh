timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -p no:cacheprovider 2>&1 | tail -3
MCF_BENCH_DEBUG=1 python bench.py --configs none --no-cpu-baseline --steps 10 2>&1 | grep -v INFO | cut -c1-700
python tools/boot_multi.py 1e8 1024 2>&1 | tail -1
tools/ab_scan.sh base 2>&1 | tail -1
