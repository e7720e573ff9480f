compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "select or score or importance" 2>&1 | tail -6; echo "rc1=$?"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "test_attention and (16-20 or 6-33 or 4-50 or 5-16 or 10-70 or 8-12 or 2-17)" 2>&1 | tail -6; echo "rc2=$?"
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "test_select or split_path" 2>&1 | tail -6; echo "rc3=$?"
