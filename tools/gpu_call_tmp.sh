timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "score or select or importance" > gpurun_out/fs_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/fs_tests.log
timeout 300 python tools/score_small_bench.py > gpurun_out/score_small_fused.txt 2>&1; cat gpurun_out/score_small_fused.txt
RAJNI_SCORE_TWO_LAUNCHES=1 timeout 300 python tools/score_small_bench.py > gpurun_out/score_small_two.txt 2>&1; cat gpurun_out/score_small_two.txt
