python -m pytest tests/test_gpu_kernels.py -x -q -k "score or select or importance" 2>&1 | tail -2
python tools/score_small_bench.py
