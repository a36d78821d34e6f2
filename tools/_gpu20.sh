timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_api.py tests/test_gpu_edges.py -x -q -p no:cacheprovider -k "boot or stats or engine" 2>&1 | tail -5
python tools/prof_boot.py 1e8 512
