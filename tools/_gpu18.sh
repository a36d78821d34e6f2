timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_api.py tests/test_gpu_edges.py -x -q -p no:cacheprovider -k "boot or stats or engine" 2>&1 | tail -15
python tools/boot_multi.py 1e8 1024 2>&1 | tail -1
