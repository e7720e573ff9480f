python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -4
python -m pytest tests/test_gpu_e2e.py -x -q 2>&1 | tail -3
python tools/config_profile.py C3 C4 2>&1 | grep -v "gemm:head\|layernorm\|im2col\|embed"
